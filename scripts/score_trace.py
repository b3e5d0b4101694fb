"""In-kernel timeline of the scoring kernel (CTA 0): prints, per role, the time spent between consecutive trace points
(wait vs work).  usage: score_trace.py [B] [NQ]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops, _lib
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 256
CTA = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
head, _, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
g = torch.Generator(device=dev).manual_seed(3)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
geo_local = rnd(B, NQ, 6)
q_h = torch.nn.functional.normalize(rnd(B, NQ, 4), dim=-1)
t_h = rnd(B, NQ, 3) * 0.3
q0 = torch.nn.functional.normalize(rnd(B, 4), dim=-1)
t0 = rnd(B, 3) * 0.3
fr, ft, fr0, ft0 = rnd(B, NQ, 256), rnd(B, NQ, 256), rnd(B, 256), rnd(B, 256)
mnum = torch.full((B,), NQ, device=dev, dtype=torch.int32)
pk = head.prepare_tc()


def once():
    ops.score_aggregate(geo_local, q_h, t_h, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"], pk["param_score_proj"],
                        head.rots.weight, head.rots.bias, head.trans.weight, head.trans.bias, want_scores=False, pack=pk["score_pack"], vecs_host=pk["score_vecs_host"])


for _ in range(3):
    once()
torch.cuda.synchronize()
buf = torch.zeros(8, 256, dtype=torch.int64, device=dev)
_lib.lib().nsac_debug_score_trace(buf.data_ptr(), CTA)
once()
torch.cuda.synchronize()
_lib.lib().nsac_debug_score_trace(None, 0)
t = buf.cpu()
t0_ = int(t[t > 0].min())
st = [int(x) - t0_ for x in t[4].tolist() if x > 0]
en = [int(x) - t0_ for x in t[5].tolist() if x > 0]
print('CTA start (ns): min', min(st), 'max', max(st))
print('CTA end   (ns):', ' '.join(str(e) for e in en))
print(f"CTA {CTA}")
names = ["residual", "mma", "epilogue", "gather", "", "", "f-producer", "f-consumer"]
for r in (0, 1, 2, 3, 6, 7):
    ev = [int(x) - t0_ for x in t[r].tolist() if x > 0]
    print(f"{names[r]:9s} n={len(ev)} first {ev[0] if ev else None} last {ev[-1] if ev else None}")
    print("   t(ns):", " ".join(str(e) for e in ev[:80]))
    print("   dt   :", " ".join(str(b - a) for a, b in zip(ev[:79], ev[1:80])))

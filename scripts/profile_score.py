"""Runs only the hypothesis-scoring call (K8+K9) on the roofline configuration — the target of
`ncu --set full -k regex:score` captures (never a timing source)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
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
for _ in range(reps):
    ops.score_aggregate(geo_local, q_h, t_h, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"], pk["param_score_proj"],
                        head.rots.weight, head.rots.bias, head.trans.weight, head.trans.bias, want_scores=False, pack=pk["score_pack"], vecs_host=pk["score_vecs_host"])
torch.cuda.synchronize()
print("done")

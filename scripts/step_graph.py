"""Is a bench step GPU-bound or launch-bound?  Times the camera-head step (bench.py workload) eagerly and as CUDA-graph
replays, and per stage (pixel pose net / matcher / hypothesis features / scoring) with CUDA events."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops, synthetic
from tests import util
import bench

dev = torch.device("cuda:0")
B, NQ, P = 64, bench.NQ, bench.PLANES
head, match, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
hp = synthetic.all_pairs_hypotheses(P, NQ).to(dev, torch.int32)
host = synthetic.make_batch(0, B, P)
f1, f2 = synthetic.device_features(B, dev, seed=7)
d = host.to(dev)


def step():
    return head(f1, f2, d.planes1, d.planes2, d.app1, d.app2, matching_net=match, hyp_pairs=hp)[5]["pose"]


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.reset_launch_count()
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    step()
e1.record()
t_enq = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
eager = e0.elapsed_time(e1) / 10
launches = ops.launch_count() / 10

# stage times (eager, events around the stages)
marks = {}
def wrap(obj, name, tag):
    fn = getattr(obj, name)
    def w(*a, **k):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); r = fn(*a, **k); a1.record()
        marks.setdefault(tag, []).append((a0, a1))
        return r
    setattr(obj, name, w)
wrap(head, "_forward_pixel_camera_head", "pixel pose net (K1)")
wrap(match, "match", "matcher (K3-K5)")
wrap(head, "_hypothesis_features", "hypothesis features (K7)")
wrap(ops, "score_aggregate", "scoring (K8+K9)")
for _ in range(3):
    step()
torch.cuda.synchronize()
stages = {k: sum(a.elapsed_time(b) for a, b in v[-3:]) / 3 for k, v in marks.items()}

out = {"eager_ms_per_step": eager, "host_enqueue_ms_per_step": t_enq * 1e3, "launches_per_step": launches, "stages_ms": stages}
try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = step()
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    out["graph_ms_per_step"] = e0.elapsed_time(e1) / 10
except Exception as e:  # noqa: BLE001
    out["graph_error"] = f"{type(e).__name__}: {e}"[:400]
print(json.dumps(out))

"""S5 step (bench.py's headline workload: uint8 RGB -> backbone -> camera head, 64 pairs) eagerly vs as CUDA-graph replays:
how much of the step is launch gaps?  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nopesac_b200 import config, meta_arch, ops, synthetic
from tests import util

dev = torch.device("cuda:0")
B, NQ, P = 64, bench.NQ, bench.PLANES
model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ, "soft", 0.2), with_backbone=True)
sd, msd = util.make_weights(NQ)
model.camera_head_list[0].load_state_dict(sd)
model.matching_head.load_state_dict(msd)
model.backbone.load_state_dict(bench.backbone_state())
model = model.to(dev)
hp = synthetic.all_pairs_hypotheses(P, NQ).to(dev, torch.int32)
d = synthetic.make_batch(0, B, P).to(dev)
images = synthetic.make_images(7000, 2 * B, 480, 640).to(dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    return model.inference_from_images(images, None, d.planes1, d.planes2, d.app1, d.app2, hyp_pairs=hp)[5]["pose"]


def timed(fn, n=10, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        if flush:
            flush_buf.fill_(1)
        fn()
    e1.record()
    enq = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, enq * 1e3


out = {}
out["eager_ms"], out["eager_host_enqueue_ms"] = timed(step)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    rows = step()
ref = step().clone()
g.replay()
torch.cuda.synchronize()
out["graph_equals_eager"] = bool(torch.equal(rows, ref))
out["graph_ms"], out["graph_host_enqueue_ms"] = timed(g.replay)
out["eager_ms_no_flush"], _ = timed(step, flush=False)
out["graph_ms_no_flush"], _ = timed(g.replay, flush=False)
print(json.dumps(out))

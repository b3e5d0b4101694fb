"""PlaneTRHead (row f1, first half) on the bench's 128 images: one marked call for an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_planetr.py [pairs]
       python scripts/summarize_launches.py out.csv        (the call after the last uint8-fill marker)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nopesac_b200 import config, meta_arch, synthetic  # noqa: E402
from tests import util  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
nq = 50
model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(nq), with_backbone=True, with_plane_head=True)
model.sem_seg_head.load_state_dict(synthetic.make_weights(util.planetr_shapes(nq), 77))
shapes = {k: tuple(v.shape) for k, v in model.backbone.state_dict().items()}
model.backbone.load_state_dict(synthetic.make_backbone_weights(shapes, seed=8))
model = model.to(dev)
images = synthetic.make_images(9100, 2 * pairs, 480, 640).to(dev)
feats = model.backbone(images, planes=True)
for _ in range(2):
    out, qf = model.sem_seg_head(feats)
torch.cuda.synchronize()
marker = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
marker.fill_(1)                                    # the summariser's step marker
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out, qf = model.sem_seg_head(feats)
e1.record()
torch.cuda.synchronize()
print(f"PlaneTRHead on {2 * pairs} images: {e0.elapsed_time(e1):.3f} ms (under a profiler this is not a measurement)")

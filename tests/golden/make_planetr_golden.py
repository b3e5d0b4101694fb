"""Generates tests/golden/planetr_nq20.golden and planetr_state_shapes.json by running the LIVE, UNMODIFIED reference PlaneTRHead
(/root/reference, via oracle/ref_planetr_loader.py) on seeded synthetic feature maps with seeded synthetic weights.  Build
container only:    python tests/golden/make_planetr_golden.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nopesac_b200 import synthetic  # noqa: E402
from oracle import ref_planetr_loader as L  # noqa: E402
from tests.test_oracle_planetr import make_features  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
head50 = L.build_head(50)
with open(os.path.join(HERE, "planetr_state_shapes.json"), "w") as f:
    json.dump({k: list(v.shape) for k, v in head50.state_dict().items()}, f, indent=0, sort_keys=True)
case = dict(NQ=20, N=2, H=96, W=128, weight_seed=77, feat_seed=3)
head = L.build_head(case["NQ"])
sd = synthetic.make_weights({k: tuple(v.shape) for k, v in head.state_dict().items()}, case["weight_seed"])
head.load_state_dict(sd)
with torch.no_grad():
    out, hs = head(make_features(case["feat_seed"], case["N"], case["H"], case["W"]))
out = {k: v.clone() for k, v in out.items()}
out["query_feat"] = hs.clone()
torch.save({"case": case, "outputs": out}, os.path.join(HERE, "planetr_nq20.golden"))
print("wrote", {k: tuple(v.shape) for k, v in out.items()})

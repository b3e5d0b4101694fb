"""Times one nsac_score_aggregate_tc call (B=512, m=NQ=256) by CUDA-graph replay, like bench.py's roofline leg; one JSON line.
Used to A/B library variants: NSAC_B200_LIB=build/variants/x.so python scripts/score_quick.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
once = bench.score_case(dev, 512, 256)
ms = bench.graph_timed(once, 30)
alg = bench.score_algorithmic_bytes(512, 256, 256)
print(json.dumps({"lib": os.environ.get("NSAC_B200_LIB", "default"), "us": ms * 1e3, "frac_of_hbm": alg / ms / 1e6 / bench.measured_peaks()["hbm"]}))

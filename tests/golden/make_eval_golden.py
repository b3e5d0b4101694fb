"""Generates tests/golden/camera_eval.json from the reference's own evaluation source (oracle/ref_eval_loader.py):
seeded predictions / ground truth and the metrics table `_eval_camera_reg` prints for them.  Run in the build container."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_eval_loader  # noqa: E402


def make_case(seed, n):
    rng = np.random.RandomState(seed)
    gt_q = rng.randn(n, 4); gt_q /= np.linalg.norm(gt_q, axis=1, keepdims=True)
    gt_q *= np.sign(gt_q[:, :1] + 1e-12)
    gt_t = rng.uniform(-1.5, 1.5, (n, 3))
    scale = rng.choice([0.02, 0.1, 0.4, 1.0], size=(n, 1))
    pr_q = gt_q + scale * 0.5 * rng.randn(n, 4); pr_q /= np.linalg.norm(pr_q, axis=1, keepdims=True)
    pr_t = gt_t + scale * rng.randn(n, 3)
    return [a.astype(np.float32) for a in (pr_t, pr_q, gt_t, gt_q)]


if __name__ == "__main__":
    assert ref_eval_loader.available(), "needs /root/reference"
    _, eval_ref = ref_eval_loader.load()
    cases = []
    for seed, n in ((1, 1), (2, 2), (3, 17), (4, 64), (5, 513)):
        pr_t, pr_q, gt_t, gt_q = make_case(seed, n)
        preds = [{"camera": {"pred": {"tran": pr_t[i], "rot": pr_q[i]}, "gts": {"tran": gt_t[i], "rot": gt_q[i]}}} for i in range(n)]
        m = eval_ref(preds, "camera")
        cases.append({"seed": seed, "n": n, "pred_tran": pr_t.tolist(), "pred_rot": pr_q.tolist(), "gt_tran": gt_t.tolist(),
                      "gt_rot": gt_q.tolist(), "metrics": {k: float(v) for k, v in m.items()}})
    with open(os.path.join(ROOT, "tests", "golden", "camera_eval.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "cases")

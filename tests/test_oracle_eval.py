"""CPU: the evaluation oracle (oracle/eval_restate.py) against the reference's own source (when /root/reference is
present) and against the committed golden fixture generated from it."""
import json
import os

import numpy as np
import pytest

from oracle import eval_restate, ref_eval_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    with open(os.path.join(ROOT, "tests", "golden", "camera_eval.json")) as f:
        return json.load(f)


def test_eval_oracle_matches_golden():
    for c in _cases():
        arr = lambda k: np.asarray(c[k], dtype=np.float32)
        m = eval_restate.eval_camera_reg(arr("pred_tran"), arr("pred_rot"), arr("gt_tran"), arr("gt_rot"))
        assert set(m) == set(c["metrics"])
        for k, v in c["metrics"].items():
            assert float(m[k]) == v, (c["n"], k, float(m[k]), v)          # bit-exact: same numpy operations


@pytest.mark.skipif(not ref_eval_loader.available(), reason="reference tree not present")
def test_eval_oracle_matches_reference_source():
    angle_ref, eval_ref = ref_eval_loader.load()
    rng = np.random.RandomState(7)
    for n in (1, 2, 5, 100):
        q1 = rng.randn(n, 4).astype(np.float32); q1 /= np.linalg.norm(q1, axis=1, keepdims=True)
        q2 = rng.randn(n, 4).astype(np.float32); q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
        t1, t2 = rng.randn(n, 3).astype(np.float32), rng.randn(n, 3).astype(np.float32)
        assert np.array_equal(angle_ref(q1, q2), eval_restate.angle_error_vec(q1, q2))
        preds = [{"camera": {"pred": {"tran": t1[i], "rot": q1[i]}, "gts": {"tran": t2[i], "rot": q2[i]}}} for i in range(n)]
        want = eval_ref(preds, "camera")
        got = eval_restate.eval_camera_reg(t1, q1, t2, q2)
        assert set(want) == set(got)
        for k in want:
            assert float(want[k]) == float(got[k]), (n, k)
    # edge: identical / antipodal quaternions -> 0 degrees (|q.q_gt| clipped to 1)
    q = np.array([[1, 0, 0, 0], [0.5, 0.5, 0.5, 0.5]], dtype=np.float32)
    assert np.allclose(eval_restate.angle_error_vec(q, -q), 0.0) and np.array_equal(angle_ref(q, -q), eval_restate.angle_error_vec(q, -q))

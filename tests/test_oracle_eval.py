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


@pytest.mark.skipif(not ref_eval_loader.available(), reason="reference tree not present")
def test_result_records_match_reference_get_optimized_dict(tmp_path):
    """Host-side result formatting (nopesac_b200.evaluation.prediction_records / optimized_dict / save_results) against
    the reference's own get_optimized_dict run on the same records; runs on CPU tensors (pure formatting, no kernels)."""
    import pickle
    import torch
    from nopesac_b200 import evaluation
    g = torch.Generator().manual_seed(3)
    B, P = 3, 5
    results, gt_t, gt_q = [], torch.randn(B, 3, generator=g), torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=1)
    for i in range(B):
        r = {"0": {"image_id": f"a{i}", "file_name": f"a{i}.png", "pred_plane": torch.randn(P, 3, generator=g)},
             "1": {"image_id": f"b{i}", "file_name": f"b{i}.png", "pred_plane": torch.randn(P, 3, generator=g)},
             "pred_aff": None, "depth": {"0": None, "1": None}}
        for key in ("camera_zero", "camera_init", "camera"):
            r[key] = {"tran": torch.randn(3, generator=g), "rot": torch.nn.functional.normalize(torch.randn(4, generator=g), dim=0)}
        r["pred_assignment"] = (torch.rand(P, P, generator=g) > 0.7).float()
        r["pred_assignment_beforeRef0"] = r["pred_assignment"].clone()
        results.append(r)
    recs = evaluation.prediction_records(results, gt_t, gt_q)
    want = ref_eval_loader.load_get_optimized_dict()(recs)
    got = evaluation.optimized_dict(recs)
    assert sorted(want) == sorted(got)

    def same(a, b):
        if isinstance(a, dict):
            assert sorted(a) == sorted(b)
            for k in a:
                same(a[k], b[k])
        elif isinstance(a, np.ndarray):
            assert np.array_equal(a, b)
        else:
            assert a == b
    for k in want:
        same(want[k], got[k])
    evaluation.save_results(recs, str(tmp_path))
    with open(tmp_path / "continuous.pkl", "rb") as f:
        disk = pickle.load(f)
    for k in want:
        same(want[k], disk[k])
    back = torch.load(tmp_path / "NopeSAC_instances_predictions.pth", weights_only=False)
    assert len(back) == B and np.array_equal(back[1]["camera"]["pred"]["tran"], recs[1]["camera"]["pred"]["tran"])
    # the records feed the reference's metric code unchanged
    _, eval_ref = ref_eval_loader.load()
    m = eval_ref(recs, "camera")
    mine = eval_restate.eval_camera_reg(np.stack([r["camera"]["pred"]["tran"] for r in recs]), np.stack([r["camera"]["pred"]["rot"] for r in recs]),
                                        gt_t.numpy(), gt_q.numpy())
    for k in m:
        assert float(m[k]) == float(mine[k])

"""TEST INFRASTRUCTURE — not product code.

CPU restatement (numpy) of the reference's camera-pose evaluation, SURVEY.md §8 row f3:
  * `angle_error_vec`            evaluation/mp3d_evaluation.py:463-465
  * `MP3DEvaluator._eval_camera_reg`  evaluation/mp3d_evaluation.py:382-425 (error vectors, thresholds, metric names)
Pinned against the reference's own source (tests/test_oracle_eval.py executes those two definitions straight from
/root/reference/NopeSAC_Net/evaluation/mp3d_evaluation.py, bit-exact) and against tests/golden/camera_eval.json, which was
generated from that source (tests/golden/make_eval_golden.py).  Only tests/, smoke() and bench.py's CPU legs may import it.
"""
from __future__ import annotations

import numpy as np


def angle_error_vec(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """mp3d_evaluation.py:463-465."""
    assert v1.ndim == 2 and v2.ndim == 2
    return 2 * np.arccos(np.clip(np.abs(np.sum(np.multiply(v1, v2), axis=1)), -1.0, 1.0)) * 180 / np.pi


def eval_camera_reg(pred_tran, pred_rot, gt_tran, gt_rot) -> dict:
    """mp3d_evaluation.py:382-425: the camera metrics table from stacked predictions [n,3] / [n,4] and ground truth."""
    gt_tran, gt_rot = np.asarray(gt_tran), np.asarray(gt_rot)
    pred_tran, pred_rot = np.asarray(pred_tran), np.asarray(pred_rot)
    err_t = np.linalg.norm(gt_tran - pred_tran, axis=1)              # :389
    err_r = angle_error_vec(pred_rot, gt_rot)                        # :390
    n = len(err_t)
    return {
        "T median err": np.median(err_t),
        "T mean err": np.mean(err_t),
        "T err < 1.0": (err_t < 1.0).sum() / n * 100,
        "T err < 0.5": (err_t < 0.5).sum() / n * 100,
        "T err < 0.2": (err_t < 0.2).sum() / n * 100,
        "R median err": np.median(err_r),
        "R mean err": np.mean(err_r),
        "R err < 30": (err_r < 30).sum() / n * 100,
        "R err < 15": (err_r < 15).sum() / n * 100,
        "R err < 10": (err_r < 10).sum() / n * 100,
    }

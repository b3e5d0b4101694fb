"""Camera-pose evaluation of the result rows — the step after the hot path.

Mirrors `MP3DEvaluator._eval_camera_reg` (evaluation/mp3d_evaluation.py:382-425): same error definitions
(`angle_error_vec`, :463-465), thresholds and metric names, so the dict can be logged / compared like the reference's
`camera metrics` table.  Errors and threshold counts are computed by `nsac_camera_errors` on the device from the `[B,16]`
result rows (the unit that is exchanged between GPUs); only the 10 scalars come back to the host.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from typing import Dict, List

import torch

from . import _lib

METRIC_KEYS = ("T median err", "T mean err", "T err < 1.0", "T err < 0.5", "T err < 0.2",
               "R median err", "R mean err", "R err < 30", "R err < 15", "R err < 10")


def _median_like_numpy(sorted_vals: torch.Tensor) -> torch.Tensor:
    """np.median: mean of the two middle values for an even count (torch.median returns the lower one)."""
    n = sorted_vals.numel()
    return sorted_vals[n // 2] if n % 2 else (sorted_vals[n // 2 - 1] + sorted_vals[n // 2]) * 0.5


def camera_errors(pose_rows: torch.Tensor, gt_tran: torch.Tensor, gt_rot: torch.Tensor):
    """pose_rows [B, >=7] (t, q, ...) CUDA fp32 -> (err_t [B], err_r [B] in degrees, stats [8]) on the device."""
    if not (pose_rows.is_cuda and gt_tran.is_cuda and gt_rot.is_cuda):
        raise RuntimeError("nopesac_b200.evaluation: tensors must live on a CUDA device (there is no CPU fallback)")
    assert pose_rows.dtype == torch.float32 and pose_rows.dim() == 2 and pose_rows.shape[1] >= 7 and pose_rows.stride(1) == 1
    B = pose_rows.shape[0]
    gt_tran = gt_tran.to(torch.float32).contiguous().view(B, 3)
    gt_rot = gt_rot.to(torch.float32).contiguous().view(B, 4)
    err_t = torch.empty(B, device=pose_rows.device)
    err_r = torch.empty(B, device=pose_rows.device)
    stats = torch.empty(8, device=pose_rows.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = _lib.lib().nsac_camera_errors(C.cast(p(pose_rows), C.POINTER(C.c_float)), pose_rows.stride(0),
                                       C.cast(p(gt_tran), C.POINTER(C.c_float)), C.cast(p(gt_rot), C.POINTER(C.c_float)), B,
                                       C.cast(p(err_t), C.POINTER(C.c_float)), C.cast(p(err_r), C.POINTER(C.c_float)),
                                       C.cast(p(stats), C.POINTER(C.c_float)),
                                       C.c_void_p(torch.cuda.current_stream(pose_rows.device).cuda_stream))
    _lib.check(st, "nsac_camera_errors")
    return err_t, err_r, stats


def camera_metrics(pose_rows: torch.Tensor, gt_tran: torch.Tensor, gt_rot: torch.Tensor) -> Dict[str, float]:
    """The reference's camera metrics table (one host read of 10 scalars)."""
    err_t, err_r, stats = camera_errors(pose_rows, gt_tran, gt_rot)
    B = pose_rows.shape[0]
    med = torch.stack([_median_like_numpy(torch.sort(err_t).values), _median_like_numpy(torch.sort(err_r).values)])
    vals = torch.cat([med, stats]).cpu().tolist()
    t_med, r_med, t_mean, r_mean, c1, c05, c02, c30, c15, c10 = vals
    pct = lambda c: c / B * 100.0
    return {"T median err": t_med, "T mean err": t_mean, "T err < 1.0": pct(c1), "T err < 0.5": pct(c05), "T err < 0.2": pct(c02),
            "R median err": r_med, "R mean err": r_mean, "R err < 30": pct(c30), "R err < 15": pct(c15), "R err < 10": pct(c10)}


class CameraEvaluator:
    """`process()` batches of result rows with their ground truth, `evaluate()` once — the shape of the reference's
    DatasetEvaluator (mp3d_evaluation.py:184-258, 259-313) for the camera task."""

    def __init__(self):
        self._rows, self._gt_t, self._gt_q = [], [], []

    def reset(self):
        self._rows, self._gt_t, self._gt_q = [], [], []

    def process(self, pose_rows: torch.Tensor, gt_tran: torch.Tensor, gt_rot: torch.Tensor):
        self._rows.append(pose_rows[:, :7].detach().clone())
        self._gt_t.append(gt_tran.detach().to(pose_rows.device, torch.float32).view(-1, 3))
        self._gt_q.append(gt_rot.detach().to(pose_rows.device, torch.float32).view(-1, 4))

    def evaluate(self) -> Dict[str, float]:
        if not self._rows:
            return {}
        return camera_metrics(torch.cat(self._rows).contiguous(), torch.cat(self._gt_t), torch.cat(self._gt_q))


# ---------------------------------------------------------------------------------------------------------------------
# Result files (mp3d_evaluation.py:259-313 `get_optimized_dict`, :336-342, :852-860 `save_dict`): the per-pair records that
# the reference's `eval.py --evaluate camera` reads back from `continuous.pkl` / `NopeSAC_instances_predictions.pth`.
# ---------------------------------------------------------------------------------------------------------------------
def prediction_records(results: List[dict], gt_tran, gt_rot) -> List[dict]:
    """`PlaneTR_NopeSAC.inference` results (one dict per pair) + ground truth -> the evaluator's prediction records
    (mp3d_evaluation.py:184-258 `process`): every `camera*` key becomes {"pred": {tran, rot}, "gts": {tran, rot}},
    assignments and per-view plane parameters move to the host.  One device->host copy per tensor, at the end."""
    recs = []
    for i, r in enumerate(results):
        gts = {"tran": _np(gt_tran[i]), "rot": _np(gt_rot[i])}
        rec = {"0": dict(r["0"]), "1": dict(r["1"])}
        for key, value in r.items():
            if "camera" in key and isinstance(value, dict) and "tran" in value:
                rec[key] = {"pred": {"tran": _np(value["tran"]), "rot": _np(value["rot"])}, "gts": gts}
            elif "assignment" in key:
                rec[key] = value.detach().cpu() if isinstance(value, torch.Tensor) else value
        recs.append(rec)
    return recs


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def optimized_dict(predictions: List[dict]) -> Dict[int, dict]:
    """mp3d_evaluation.py:259-313: the `continuous.pkl` payload."""
    out = {}
    for idx, prediction in enumerate(predictions):
        best_assignment = prediction["pred_assignment"].numpy()
        camera = prediction["camera"]
        aux = {key: {"position": prediction[key]["pred"]["tran"], "rotation": prediction[key]["pred"]["rot"]}
               for key in prediction if "camera" in key}
        del aux          # the reference builds it and drops it too (:274-280)
        out[idx] = {
            "n_corr": best_assignment.sum(),
            "cost": 0.1,
            "best_camera": {"position": camera["pred"]["tran"], "rotation": camera["pred"]["rot"]},
            "gt_camera": {"position": camera["gts"]["tran"], "rotation": camera["gts"]["rot"]},
            "best_assignment": best_assignment,
            "plane_param_override": {"0": _np(prediction["0"]["pred_plane"]), "1": _np(prediction["1"]["pred_plane"])},
            "image_ids": {"0": prediction["0"]["image_id"], "1": prediction["1"]["image_id"]},
        }
    return out


def save_results(predictions: List[dict], output_dir: str):
    """Writes `NopeSAC_instances_predictions.pth` and `continuous.pkl` like MP3DEvaluator.evaluate (:331-342)."""
    os.makedirs(output_dir, exist_ok=True)
    torch.save(predictions, os.path.join(output_dir, "NopeSAC_instances_predictions.pth"))
    with open(os.path.join(output_dir, "continuous.pkl"), "wb") as f:
        pickle.dump(optimized_dict(predictions).copy(), f)

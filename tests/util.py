"""Shared helpers of the test-suite: golden fixtures, seeded inputs/weights, oracle and CUDA runners."""
from __future__ import annotations

import glob
import json
import os

import torch

from nopesac_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances of the parity bar (BASELINE.json north_star / SURVEY.md §7, §8d)
ABS_TOL = 1e-4          # poses, scores, exp(log_scores_padded)
LOG_REL_TOL = 1e-4      # raw log-scores, relative


def golden_names():
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def state_shapes(num_queries: int):
    with open(os.path.join(GOLDEN_DIR, "state_shapes.json")) as f:
        all_shapes = json.load(f)
    s = all_shapes[str(num_queries)]
    return s["head"], s["match"]


def head_shapes_for(num_queries: int):
    """Shapes for any NQ (only the two score-MLP input layers depend on it)."""
    head, match = state_shapes(50)
    head = dict(head)
    for k in ("normal_score_proj.layers.0.weight", "param_score_proj.layers.0.weight"):
        head[k] = [128, num_queries]
    return head, match


def make_weights(num_queries: int, head_seed=40, match_seed=41):
    hs, ms = head_shapes_for(num_queries)
    return synthetic.make_weights(hs, head_seed), synthetic.make_weights(ms, match_seed)


def initial_pose_for(pair_idx: int):
    """Must stay identical to tests/golden/make_golden.py:initial_pose_for."""
    g = torch.Generator().manual_seed(9000 + pair_idx)
    q = torch.nn.functional.normalize(torch.randn(1, 4, generator=g), dim=-1)
    t = (torch.rand(1, 3, generator=g) * 2 - 1) * 0.5
    return t, q


def case_batch(case, pair_idx):
    b = synthetic.make_batch(pair_idx, 1, case["P"], with_features=case["feats"], negative_k=case["negk"])
    if "P1" in case:
        b.planes1, b.app1 = b.planes1[:, :case["P1"]].contiguous(), b.app1[:, :case["P1"]].contiguous()
    return b


def case_hyp_pairs(case):
    return None if case["hyp"] == "match" else synthetic.all_pairs_hypotheses(case["P"], case["hyp"])


def run_oracle_case(case, pair_idx, sd, msd):
    from oracle import restate
    b = case_batch(case, pair_idx)
    ip = None if case["feats"] else initial_pose_for(pair_idx)
    with torch.no_grad():
        return restate.inference_joint(sd, msd, b.feats1, b.feats2, b.planes1, b.planes2, b.app1, b.app2,
                                       num_queries=case["NQ"], out_cam_type=case["cam"], match_threshold=case["thr"],
                                       hyp_pairs=case_hyp_pairs(case), initial_pose=ip)


def oracle_to_flat(o):
    """restate.inference_joint output -> the flat key layout of the golden fixtures."""
    r = o["ref"]
    f = {
        "camera_init_t": o["camera_init"][0], "camera_init_q": o["camera_init"][1],
        "camera_initRec_t": o["camera_initRec"][0], "camera_initRec_q": o["camera_initRec"][1],
        "log_scores_padded": o["log_scores_padded"], "assignment_before": o["assignment_before"],
        "assignment_after": o["assignment_after"],
        "geo_local": o["geo_local"], "geo_global": o["geo_global"], "sig_seq": o["sig_seq"][:, 0],
        "matched_num": torch.tensor(o["matched_num"]),
    }
    for k in ("pred_trans", "pred_rot", "pred_trans_avg", "pred_rot_avg", "all_pred_trans", "all_pred_rots",
              "score_soft_rot", "score_soft_offset", "l2_dist", "normal_dist", "offset_dist"):
        if k in r:
            f[k] = r[k]
    return f


def build_cuda_heads(num_queries, out_cam_type="soft", match_threshold=0.2, device="cuda"):
    from nopesac_b200 import config
    from nopesac_b200.camera_head import build_camera_head
    from nopesac_b200.matching_head import build_matching_head
    from nopesac_b200.meta_arch import RESNET50_OUTPUT_SHAPE
    cfg = config.inference_cfg(num_queries, out_cam_type, match_threshold)
    head = build_camera_head(cfg, RESNET50_OUTPUT_SHAPE).eval()
    match = build_matching_head(cfg).eval()
    sd, msd = make_weights(num_queries)
    head.load_state_dict(sd)
    match.load_state_dict(msd)
    return head.to(device), match.to(device), sd, msd


def maxdiff(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) if a.numel() else 0.0


def planetr_shapes(num_queries: int):
    """State-dict shapes of the reference's PlaneTRHead (tests/golden/planetr_state_shapes.json, written from the live reference
    module at NUM_OBJECT_QUERIES = 50; only `query_embed.weight` depends on it)."""
    with open(os.path.join(GOLDEN_DIR, "planetr_state_shapes.json")) as f:
        shapes = json.load(f)
    shapes["query_embed.weight"] = [num_queries, shapes["query_embed.weight"][1]]
    return shapes

"""Generates the committed golden fixtures (tests/golden/*.pt) by running the LIVE, UNMODIFIED reference
code from /root/reference (via oracle/ref_loader.py) on seeded synthetic pairs with seeded synthetic
weights.  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the case description (so inputs and weights can be regenerated bit-identically from
nopesac_b200.synthetic.make_batch / make_weights) and the reference outputs.  Cases whose hypothesis list is
not the matcher's own assignment (the "P planes x H hypotheses" stress mapping of SURVEY.md §8(d)) call the
reference's own stage methods in the order of inference_Joint (camera_head.py:433-583).
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nopesac_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HEAD_SEED, MATCH_SEED = 40, 41

# name, NQ, P, pairs, features(K1), hyp ("match" = matcher's assignment, int = all-pairs list of that length),
# cam type, match threshold, negative_k
CASES = [
    dict(name="c1_nq32_p8", NQ=32, P=8, pairs=[0, 1], feats=True, hyp="match", cam="soft", thr=0.2, negk=False),
    dict(name="mp3d_nq50_p16", NQ=50, P=16, pairs=[0, 1, 2, 3], feats=True, hyp="match", cam="soft", thr=0.2, negk=False),
    dict(name="mp3d_nq50_p16_nofeat", NQ=50, P=16, pairs=[4, 5, 6, 7, 8, 9], feats=False, hyp="match", cam="soft", thr=0.2, negk=False),
    dict(name="mincost_nq50_p16", NQ=50, P=16, pairs=[0, 1, 2], feats=False, hyp="match", cam="min-cost", thr=0.2, negk=False),
    dict(name="maxscore_nq50_p16", NQ=50, P=16, pairs=[0, 1, 2], feats=False, hyp="match", cam="max-score", thr=0.2, negk=False),
    dict(name="avgall_nq50_p16", NQ=50, P=16, pairs=[0, 1], feats=False, hyp="match", cam="avg-all", thr=0.2, negk=False),
    dict(name="negk_nq50_p16", NQ=50, P=16, pairs=[0, 1, 2], feats=False, hyp="match", cam="soft", thr=0.2, negk=True),
    dict(name="nomatch_nq50_p16", NQ=50, P=16, pairs=[0, 1], feats=False, hyp="match", cam="soft", thr=1.5, negk=False),
    dict(name="onematch_nq50_p16", NQ=50, P=16, pairs=[0, 1], feats=False, hyp=1, cam="soft", thr=0.2, negk=False),
    dict(name="ragged_nq50_p5x9", NQ=50, P=9, P1=5, pairs=[0, 1], feats=False, hyp="match", cam="soft", thr=0.2, negk=False),
    dict(name="stress_nq256_p16", NQ=256, P=16, pairs=[0, 1, 2], feats=False, hyp=256, cam="soft", thr=0.2, negk=False),
    dict(name="stress_nq256_p16_feat", NQ=256, P=16, pairs=[3], feats=True, hyp=256, cam="soft", thr=0.2, negk=False),
    dict(name="stress_nq128_p16", NQ=128, P=16, pairs=[0, 1], feats=False, hyp=128, cam="min-cost", thr=0.2, negk=False),
    dict(name="stress_nq512_p16", NQ=512, P=16, pairs=[0], feats=False, hyp=512, cam="max-score", thr=0.2, negk=False),
    dict(name="partial_nq64_p16", NQ=64, P=16, pairs=[0, 1], feats=False, hyp=40, cam="soft", thr=0.2, negk=False),
]


def initial_pose_for(pair_idx: int):
    """Seeded stand-in for the pixel network's output when K1 is skipped (stage set S3)."""
    g = torch.Generator().manual_seed(9000 + pair_idx)
    q = torch.nn.functional.normalize(torch.randn(1, 4, generator=g), dim=-1)
    t = (torch.rand(1, 3, generator=g) * 2 - 1) * 0.5
    return t, q


def run_reference(head, match, b, case, pair_idx):
    """inference_Joint's own sequence with optional overrides, calling reference methods only."""
    P = ref_loader.private
    NQ = case["NQ"]
    out = {}
    with ref_loader.cpu_patch(), torch.no_grad():
        if case["feats"]:
            _, cam0, _ = P(head, "forward_PixelCameraHead")(b.feats1, b.feats2)
            t_init, q_init = cam0["pred_trans"], cam0["pred_rot"]
        else:
            t_init, q_init = initial_pose_for(pair_idx)
        if q_init[0, 0] < 0:
            q_init = -q_init
        out["camera_init_t"], out["camera_init_q"] = t_init, q_init
        _, q0, rot_feat0 = P(head, "forward_RotRecHead")(q_init)
        _, t0, trans_feat0 = P(head, "forward_TransRecHead")(t_init)
        out["camera_initRec_t"], out["camera_initRec_q"] = t0, q0
        cam = torch.cat([t0, q0], dim=-1)
        _, lsp = match(b.app1, b.app2, cam, b.planes1, b.planes2, gt_corr_matrix=None, normal_decay=1.0, offset_deacy=1.0)
        out["log_scores_padded"] = lsp
        assign = ref_loader.load().camera_modules.get_assignment_matrix(lsp, match_threshold=case["thr"])
        out["assignment_before"] = assign
        if case["hyp"] == "match":
            p1, p2, A = b.planes1, b.planes2, assign
        else:
            hp = synthetic.all_pairs_hypotheses(case["P"], case["hyp"])
            p1, p2 = b.planes1[:, hp[:, 0]], b.planes2[:, hp[:, 1]]
            A = torch.eye(hp.shape[0])[None]          # nonzero() of the identity lists the pairs in order
        dev = torch.device("cpu")
        geo_local, _, _ = head.get_pred_geo_sequence(planes1=p1, planes2=p2, pred_assignment_matrix=A, pred_cams=None, device=dev)
        cam_in = {"tran": t0, "rot": q0}
        geo_global, score_seq, mnums = head.get_pred_geo_sequence(planes1=p1, planes2=p2, pred_assignment_matrix=A, pred_cams=cam_in, device=dev)
        cam_in2 = {"tran": torch.zeros_like(t0), "rot": q0}
        geo_aux, _, _ = head.get_pred_geo_sequence(planes1=p1, planes2=p2, pred_assignment_matrix=A, pred_cams=cam_in2, device=dev)
        sig = ((geo_global[:, :, 0:1] * geo_aux[:, :, 0:1]) >= 0).float()
        sig = (sig - 0.5) * 2.
        out.update(geo_local=geo_local[0], geo_global=geo_global[0], sig_seq=sig[0, :, 0],
                   matched_num=torch.tensor(mnums[0]))
        _, r = P(head, "inference_PlaneCamRefHead")(
            trans_feat0, rot_feat0, geo_global, score_seq, gt_pose=None, geo_sequence_local=geo_local,
            matched_nums=mnums, out_cam_type=case["cam"], sig_seq=sig, initial_rot=q0, initial_trans=t0)
        for k in ("pred_trans", "pred_rot", "pred_trans_avg", "pred_rot_avg", "all_pred_trans", "all_pred_rots",
                  "score_soft_rot", "score_soft_offset"):
            if k in r:
                out[k] = r[k]
        if case["NQ"] <= 64:
            for k in ("l2_dist", "normal_dist", "offset_dist"):
                if k in r:
                    out[k] = r[k]
        # the full reference forward as an end-to-end cross-check where no override is active
        if case["hyp"] == "match" and case["feats"] and case["thr"] == 0.2:
            cams, _, _, _, ass, _ = head(b.feats1, b.feats2, b.planes1, b.planes2, b.app1, b.app2, matching_net=match)
            assert torch.equal(cams["camera"]["tran"], r["pred_trans"]) and torch.equal(cams["camera"]["rot"], r["pred_rot"])
            out["assignment_after"] = ass["pred_assignment"]
    return {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    shapes = {}
    for case in CASES:
        head, match, _ = ref_loader.build_heads(num_queries=case["NQ"])
        hshape = {k: tuple(v.shape) for k, v in head.state_dict().items()}
        mshape = {k: tuple(v.shape) for k, v in match.state_dict().items()}
        shapes[str(case["NQ"])] = {"head": hshape, "match": mshape}
        head.load_state_dict(synthetic.make_weights(hshape, HEAD_SEED))
        match.load_state_dict(synthetic.make_weights(mshape, MATCH_SEED))
        if case["thr"] != 0.2:
            head.matching_score_threshold = case["thr"]
        outs = []
        for pi in case["pairs"]:
            b = synthetic.make_batch(pi, 1, case["P"], with_features=case["feats"], negative_k=case["negk"])
            if "P1" in case:   # ragged: fewer planes in view 1 than in view 2
                b.planes1, b.app1 = b.planes1[:, :case["P1"]].contiguous(), b.app1[:, :case["P1"]].contiguous()
            outs.append(run_reference(head, match, b, case, pi))
        torch.save({"case": case, "head_seed": HEAD_SEED, "match_seed": MATCH_SEED, "outputs": outs},
                   os.path.join(HERE, case["name"] + ".pt"))
        m = [int(o["matched_num"]) for o in outs]
        print(f"{case['name']}: pairs={case['pairs']} matched={m}")
    with open(os.path.join(HERE, "state_shapes.json"), "w") as f:
        json.dump(shapes, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()

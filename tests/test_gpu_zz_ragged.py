"""GPU: ragged batches (pairs with different plane counts in one padded batch, `plane_count1/2`) and the sync-free hand-off
from row f1's plane lists to the camera head.  Named to run last: the same Python glue and the plain-SIMT kernels involved
(attention key counts, per-pair transport problems) were executed on the host against the oracle
(tests/test_host_head_glue.py, tests/test_simt_host_kernels.py); these tests add the real tensor-core engine underneath."""
import pytest
import torch

from tests import util
from tests.test_gpu_parity import _check_against

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]   # code that has not met a GPU yet: never hang the run


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_ragged_batch_matches_per_pair_oracle():
    dev = _gpu()
    from oracle import restate
    from nopesac_b200 import synthetic
    NQ, P = 50, 16
    counts = [(16, 16), (5, 9), (9, 5), (1, 7), (12, 16)]
    B = len(counts)
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    b = synthetic.make_batch(60, B, P)
    poses = [util.initial_pose_for(200 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    p1, p2, a1, a2 = b.planes1.clone(), b.planes2.clone(), b.app1.clone(), b.app2.clone()
    g = torch.Generator().manual_seed(3)
    outs = []
    with torch.no_grad():
        for i, (n1, n2) in enumerate(counts):
            outs.append(restate.inference_joint(sd, msd, None, None, p1[i:i + 1, :n1], p2[i:i + 1, :n2], a1[i:i + 1, :n1], a2[i:i + 1, :n2],
                                                num_queries=NQ, initial_pose=(ip[0][i:i + 1], ip[1][i:i + 1])))
            # padding = finite garbage (wrong planes, wrong appearance): it must not reach any valid output
            p1[i, n1:], p2[i, n2:] = 50 * torch.randn(P - n1, 3, generator=g), 50 * torch.randn(P - n2, 3, generator=g)
            a1[i, n1:], a2[i, n2:] = 3 * torch.randn(P - n1, 256, generator=g), 3 * torch.randn(P - n2, 256, generator=g)
    c1 = torch.tensor([c[0] for c in counts], dtype=torch.int32, device=dev)
    c2 = torch.tensor([c[1] for c in counts], dtype=torch.int32, device=dev)
    cams, _, _, lsp, ass, pro = head(None, None, p1.to(dev), p2.to(dev), a1.to(dev), a2.to(dev), matching_net=match,
                                     initial_pose=(ip[0].to(dev), ip[1].to(dev)), plane_count1=c1, plane_count2=c2)
    torch.cuda.synchronize()
    for i, ((n1, n2), o) in enumerate(zip(counts, outs)):
        cams_i = {k: {"tran": v["tran"][i:i + 1], "rot": v["rot"][i:i + 1]} for k, v in cams.items() if v["tran"].shape[0] == B}
        ass_i = {k: v[i:i + 1, :n1, :n2] for k, v in ass.items()}
        pro_i = {k: v[i:i + 1] for k, v in pro.items() if torch.is_tensor(v) and v.shape[0] == B}
        _check_against(util.oracle_to_flat(o), cams_i, [lsp[0][i:i + 1, :n1 + 1, :n2 + 1]], ass_i, pro_i, 0, f"ragged pair {i} ({n1}x{n2})")
        for k, v in ass.items():
            assert float(v[i, n1:].abs().sum()) == 0.0 and float(v[i, :, n2:].abs().sum()) == 0.0, (i, k)
        assert bool(torch.isinf(lsp[0][i, n1 + 1:]).all()) and bool(torch.isinf(lsp[0][i, :, n2 + 1:]).all())


def test_plane_lists_feed_the_camera_head_without_host_round_trip():
    """PlaneTRHead outputs of both views -> `inference_from_plane_heads` (plane lists + camera head on the padded lists with
    counts) == the camera head called pair by pair on the un-padded lists of the oracle's post-processing."""
    dev = _gpu()
    from nopesac_b200 import config, meta_arch, synthetic
    from oracle import planes_restate
    NQ = 50
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ)).to(dev)
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    B = 3
    v1 = synthetic.make_plane_head_batch(700, B, cases=("regular",))
    v2 = synthetic.make_plane_head_batch(800, B, cases=("regular", "regular", "zero"))
    poses = [util.initial_pose_for(300 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
    outs = lambda v: {k: v[k].to(dev) for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    (cams, _, _, lsp, ass, pro), l1, l2 = model.inference_from_plane_heads(outs(v1), v1["query_feat"].to(dev), outs(v2), v2["query_feat"].to(dev),
                                                                           None, None, max_planes=20, initial_pose=ip)
    o1 = planes_restate.postprocess_plane_head_mask(v1["pred_logits"], v1["pred_params"], v1["pred_mask_logits"], v1["query_feat"], 480, 640)
    o2 = planes_restate.postprocess_plane_head_mask(v2["pred_logits"], v2["pred_params"], v2["pred_mask_logits"], v2["query_feat"], 480, 640)
    assert l1.count.cpu().tolist() == [len(o["pred_plane_oriIdxs"]) for o in o1]
    assert l2.count.cpu().tolist() == [len(o["pred_plane_oriIdxs"]) for o in o2]
    head = model.camera_head_list[0]
    for i in range(B):
        n1, n2 = len(o1[i]["pred_plane_oriIdxs"]), len(o2[i]["pred_plane_oriIdxs"])
        one = head(None, None, o1[i]["pred_plane"][None].to(dev), o2[i]["pred_plane"][None].to(dev),
                   planeApp1=o1[i]["pred_plane_feats"].to(dev), planeApp2=o2[i]["pred_plane_feats"].to(dev),
                   matching_net=model.matching_head, initial_pose=(ip[0][i:i + 1], ip[1][i:i + 1]))
        assert torch.equal(ass["pred_assignment"][i, :n1, :n2], one[4]["pred_assignment"][0]), i
        assert util.maxdiff(cams["camera"]["tran"][i], one[0]["camera"]["tran"][0]) <= util.ABS_TOL, i
        assert util.maxdiff(cams["camera"]["rot"][i], one[0]["camera"]["rot"][0]) <= util.ABS_TOL, i
        assert int(pro["matched_num"][i]) == int(one[5]["matched_num"][0])


def test_meta_arch_inference_accepts_pairs_with_different_plane_counts():
    """`PlaneTR_NopeSAC.inference` (the reference's call, siamese_planeTR.py:338-450) on a batch whose pairs have different
    plane counts == the same model called pair by pair (the only way the reference can be called)."""
    dev = _gpu()
    from nopesac_b200 import config, meta_arch, synthetic
    NQ = 50
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ)).to(dev)
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    counts = [(16, 16), (6, 11), (9, 4)]
    b = synthetic.make_batch(5, len(counts), 16, with_features=True)
    bis = []
    for i, (n1, n2) in enumerate(counts):
        view = lambda planes, app, feats, n: {"pred_plane": planes[i, :n], "pred_plane_feats": app[i:i + 1, :n],
                                             "cam_feats": {k: v[i:i + 1] for k, v in feats.items()}}
        bis.append({"0": view(b.planes1, b.app1, b.feats1, n1), "1": view(b.planes2, b.app2, b.feats2, n2)})
    batched = model(bis)
    torch.cuda.synchronize()
    for i, (n1, n2) in enumerate(counts):
        single = model([bis[i]])[0]
        assert batched[i]["pred_assignment"].shape == (n1, n2)
        for key in ("pred_assignment_beforeRef0", "pred_assignment_afterRef0", "pred_assignment"):
            assert torch.equal(batched[i][key], single[key]), (i, key)
        for key in ("camera_init", "camera_initRec", "camera_avgRef0", "camera_softRef0", "camera"):
            assert util.maxdiff(batched[i][key]["tran"], single[key]["tran"]) <= util.ABS_TOL, (i, key)
            assert util.maxdiff(batched[i][key]["rot"], single[key]["rot"]) <= util.ABS_TOL, (i, key)

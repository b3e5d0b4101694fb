"""GPU parity at BASELINE.json's own sizes ("parity holes"):
  * configs[1] exactly: ONE batch of 64 pairs, 16 planes/view, NUM_OBJECT_QUERIES = 256 with all 16x16 plane pairs as
    hypotheses, against the per-pair oracle loop (the reference only ever runs bs = 1) — S3 for all 64 pairs, S4 (pixel
    pose network included) for a sub-batch;
  * configs[4]'s largest point: the scoring kernel at NQ = m = 2048 against the oracle's score_and_select;
  * 'max-score' selection (argmax of the scores: exact-fp32 scoring path) equals the oracle's index.
Bars as everywhere: indices / assignments / matched_num bit-exact, poses and scores 1e-4 abs."""
import pytest
import torch

from tests import util
from tests.test_gpu_parity import _check_against, _gpu

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200, method="thread")]


def _oracle_pairs(sd, msd, b, ip, NQ, hp, feats=False, cam="soft"):
    from oracle import restate
    outs = []
    with torch.no_grad():
        for i in range(b.planes1.shape[0]):
            f1 = {k: v[i:i + 1] for k, v in b.feats1.items()} if feats else None
            f2 = {k: v[i:i + 1] for k, v in b.feats2.items()} if feats else None
            outs.append(restate.inference_joint(sd, msd, f1, f2, b.planes1[i:i + 1], b.planes2[i:i + 1], b.app1[i:i + 1],
                                                b.app2[i:i + 1], num_queries=NQ, hyp_pairs=hp, out_cam_type=cam,
                                                initial_pose=None if feats else (ip[0][i:i + 1], ip[1][i:i + 1])))
    return outs


def test_config1_batch64_nq256_equals_per_pair_oracle():
    """configs[1]: B = 64, P = 16, NQ = 256, 257 hypotheses x 256 residual columns per pair — one CUDA batch vs 64 oracle calls."""
    dev = _gpu()
    from nopesac_b200 import synthetic
    NQ, P, B = 256, 16, 64
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    hp = synthetic.all_pairs_hypotheses(P, NQ)
    b = synthetic.make_batch(4000, B, P)
    poses = [util.initial_pose_for(4000 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    outs = _oracle_pairs(sd, msd, b, ip, NQ, hp)
    bd = b.to(dev)
    cams, _, _, lsp, ass, pro = head(None, None, bd.planes1, bd.planes2, bd.app1, bd.app2, matching_net=match,
                                     hyp_pairs=hp.to(dev, torch.int32), initial_pose=(ip[0].to(dev), ip[1].to(dev)))
    torch.cuda.synchronize()
    assert pro["matched_num"].cpu().tolist() == [NQ] * B
    for i, o in enumerate(outs):
        _check_against(util.oracle_to_flat(o), cams, lsp, ass, pro, i, f"configs[1] pair {i}")


def test_config1_with_pixel_network_nq256():
    """Same shape from backbone feature maps (stage set S4), 6 pairs in one batch."""
    dev = _gpu()
    from nopesac_b200 import synthetic
    NQ, P, B = 256, 16, 6
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    hp = synthetic.all_pairs_hypotheses(P, NQ)
    b = synthetic.make_batch(4100, B, P, with_features=True)
    outs = _oracle_pairs(sd, msd, b, None, NQ, hp, feats=True)
    bd = b.to(dev)
    cams, _, _, lsp, ass, pro = head(bd.feats1, bd.feats2, bd.planes1, bd.planes2, bd.app1, bd.app2, matching_net=match,
                                     hyp_pairs=hp.to(dev, torch.int32))
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        _check_against(util.oracle_to_flat(o), cams, lsp, ass, pro, i, f"configs[1]+K1 pair {i}")


@pytest.mark.parametrize("NQ,m", [(2048, 2048), (2048, 1500)])
def test_scoring_kernel_nq2048(NQ, m):
    """configs[4], largest sweep point: K8+K9 (tensor-core path) at NQ = 2048 against the oracle on identical inputs."""
    dev = _gpu()
    from nopesac_b200 import ops
    from oracle import restate
    sd, _ = util.make_weights(NQ)
    g = torch.Generator().manual_seed(77 + m)
    rnd = lambda *s: torch.randn(*s, generator=g)
    geo = rnd(1, NQ, 6)
    geo[:, m:] = 0
    fr, ft = rnd(NQ, 256) * 0.3, rnd(NQ, 256) * 0.3
    rf0, tf0 = rnd(1, 256) * 0.3, rnd(1, 256) * 0.3
    q0 = torch.nn.functional.normalize(rnd(1, 4), dim=-1)
    t0 = rnd(1, 3) * 0.3
    with torch.no_grad():
        qh = torch.nn.functional.normalize(restate.linear(sd, "rots", fr), dim=-1)
        th = restate.linear(sd, "trans", ft)
    c = lambda x: x.to(dev).contiguous()
    mlp = lambda p, r: tuple(c(sd[k]) for k in (f"{p}.layers.0.weight", f"{p}.layers.0.bias", f"{p}.layers.1.weight",
                                                f"{p}.layers.1.bias", f"{p}.layers.2.weight", f"{p}.layers.2.bias",
                                                f"{r}.weight", f"{r}.bias"))
    for cam in ("soft", "min-cost"):
        with torch.no_grad():
            want = restate.score_and_select(sd, fr, ft, rf0, tf0, geo, m, q0, t0, cam)
        res = ops.score_aggregate(c(geo), c(qh[None]), c(th[None]), c(q0), c(t0), c(fr[None]), c(ft[None]), c(rf0), c(tf0),
                                  torch.tensor([m], dtype=torch.int32, device=dev), mlp("normal_score_proj", "rot_score_reg"),
                                  mlp("param_score_proj", "trans_score_reg"), c(sd["rots.weight"]), c(sd["rots.bias"]),
                                  c(sd["trans.weight"]), c(sd["trans.bias"]), out_cam_type=cam, precision="fp16")
        torch.cuda.synchronize()
        pose = res["pose"][0].cpu()
        d = util.maxdiff
        assert int(pose[14]) == m
        assert d(res["score_rot"][0, :m + 1], want["score_soft_rot"][0, :, 0]) <= util.ABS_TOL, cam
        assert d(res["score_tran"][0, :m + 1], want["score_soft_offset"][0, :, 0]) <= util.ABS_TOL, cam
        assert float(res["score_rot"][0, m + 1:].abs().max()) == 0.0 if m < NQ else True
        assert d(pose[0:3], want["pred_trans"][0]) <= util.ABS_TOL and d(pose[3:7], want["pred_rot"][0]) <= util.ABS_TOL, cam
        assert d(pose[7:10], want["pred_trans_avg"][0]) <= util.ABS_TOL and d(pose[10:14], want["pred_rot_avg"][0]) <= util.ABS_TOL, cam
        if cam == "min-cost":
            assert res["sel_idx"][0].cpu().tolist() == [want["sel_rot"], want["sel_tran"]], cam


def test_max_score_selection_is_exact_through_the_head():
    """INFERENCE_OUT_CAM_TYPE = 'max-score' through the public head call: argmax index == oracle (the head routes this mode to
    the exact-fp32 scoring kernels), at the stress shape NQ = 256 and on matcher-fed hypotheses."""
    dev = _gpu()
    from nopesac_b200 import synthetic
    for NQ, hyp, B in ((256, 256, 4), (50, None, 4)):
        head, match, sd, msd = util.build_cuda_heads(NQ, "max-score", 0.2, dev)
        hp = None if hyp is None else synthetic.all_pairs_hypotheses(16, hyp)
        b = synthetic.make_batch(4200, B, 16)
        poses = [util.initial_pose_for(4200 + i) for i in range(B)]
        ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
        outs = _oracle_pairs(sd, msd, b, ip, NQ, hp, cam="max-score")
        bd = b.to(dev)
        cams, _, _, lsp, ass, pro = head(None, None, bd.planes1, bd.planes2, bd.app1, bd.app2, matching_net=match,
                                         hyp_pairs=None if hp is None else hp.to(dev, torch.int32),
                                         initial_pose=(ip[0].to(dev), ip[1].to(dev)))
        torch.cuda.synchronize()
        for i, o in enumerate(outs):
            if o["matched_num"] > 1:
                assert pro["sel_idx"][i].cpu().tolist() == [o["ref"]["sel_rot"], o["ref"]["sel_tran"]], (NQ, i)
            _check_against(util.oracle_to_flat(o), cams, lsp, ass, pro, i, f"max-score NQ={NQ} pair {i}")


def test_scoring_roofline_configuration_b512_nq256():
    """The exact configuration `roofline` in the bench line is measured on (B = 512 pairs, m = NQ = 256: 276.8 MB of algorithmic
    traffic): tensor-core path vs the exact-fp32 CUDA-core twin on ALL 512 pairs (1e-4), two calls give identical bytes,
    scores are a softmax over the 257 hypotheses of every pair (sum = 1), result rows carry m = 256 and unit quaternions."""
    dev = _gpu()
    from nopesac_b200 import ops
    B, NQ = 512, 256
    head, _, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    pk = head.prepare_tc()
    g = torch.Generator(device=dev).manual_seed(512)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    geo = rnd(B, NQ, 6)
    qh = torch.nn.functional.normalize(rnd(B, NQ, 4), dim=-1)
    th, q0, t0 = rnd(B, NQ, 3) * 0.3, torch.nn.functional.normalize(rnd(B, 4), dim=-1), rnd(B, 3) * 0.3
    fr, ft, fr0, ft0 = rnd(B, NQ, 256) * 0.3, rnd(B, NQ, 256) * 0.3, rnd(B, 256) * 0.3, rnd(B, 256) * 0.3
    mnum = torch.full((B,), NQ, device=dev, dtype=torch.int32)
    call = lambda precision: ops.score_aggregate(geo, qh, th, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"], pk["param_score_proj"],
                                                 head.rots.weight, head.rots.bias, head.trans.weight, head.trans.bias, out_cam_type="soft",
                                                 precision=precision, pack=pk["score_pack"], vecs_host=pk["score_vecs_host"])
    a, b, b2 = call("fp32"), call("fp16"), call("fp16")
    torch.cuda.synchronize()
    for k in ("pose", "score_rot", "score_tran"):
        assert torch.equal(b[k], b2[k]), f"{k}: two calls differ"
    assert util.maxdiff(a["score_rot"], b["score_rot"]) <= 1e-4 and util.maxdiff(a["score_tran"], b["score_tran"]) <= 1e-4
    lin = [0, 1, 2, 7, 8, 9, 10, 11, 12, 13]
    assert util.maxdiff(a["pose"][:, lin], b["pose"][:, lin]) <= 1e-4
    assert util.maxdiff(b["score_rot"].sum(1), torch.ones(B)) <= 1e-5 and util.maxdiff(b["score_tran"].sum(1), torch.ones(B)) <= 1e-5
    assert torch.equal(b["pose"][:, 14], mnum.float())
    assert util.maxdiff(b["pose"][:, 3:7].norm(dim=-1), torch.ones(B)) <= 1e-5

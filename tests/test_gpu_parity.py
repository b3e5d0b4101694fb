"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (nopesac_b200.ops -> ctypes ->
libnopesac_b200.so), against (1) the committed golden fixtures produced by the live reference and (2) the CPU
oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  bit-exact : assignment matrices, matched_num, sig_seq, argmin/argmax selections
  1e-4 abs  : every camera* tran/rot, per-hypothesis poses, score_soft_*, exp(log_scores_padded)
  1e-4 rel  : raw log-scores
"""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

TOL = util.ABS_TOL


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _run_cuda_case(case, pair_indices, want_diag=False):
    dev = _gpu()
    head, match, sd, msd = util.build_cuda_heads(case["NQ"], case["cam"], case["thr"], dev)
    bs = [util.case_batch(case, pi) for pi in pair_indices]
    cat = lambda f: torch.cat([f(b) for b in bs], 0).to(dev)
    p1, p2, a1, a2 = cat(lambda b: b.planes1), cat(lambda b: b.planes2), cat(lambda b: b.app1), cat(lambda b: b.app2)
    f1 = f2 = ip = None
    if case["feats"]:
        f1 = {k: torch.cat([b.feats1[k] for b in bs], 0).to(dev) for k in bs[0].feats1}
        f2 = {k: torch.cat([b.feats2[k] for b in bs], 0).to(dev) for k in bs[0].feats2}
    else:
        poses = [util.initial_pose_for(pi) for pi in pair_indices]
        ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
    hp = util.case_hyp_pairs(case)
    hp = None if hp is None else hp.to(dev, torch.int32)
    out = head(f1, f2, p1, p2, a1, a2, matching_net=match, hyp_pairs=hp, initial_pose=ip, want_diag=want_diag)
    torch.cuda.synchronize()
    return out


def _check_against(want, cams, lsp, ass, pro, i, tag):
    """want: flat golden/oracle dict of pair i of the batch."""
    m = int(want["matched_num"])
    assert int(pro["matched_num"][i]) == m, f"{tag}: matched_num"
    d = util.maxdiff
    assert d(cams["camera_init"]["tran"][i], want["camera_init_t"][0]) <= TOL, f"{tag}: camera_init tran"
    assert d(cams["camera_init"]["rot"][i], want["camera_init_q"][0]) <= TOL, f"{tag}: camera_init rot"
    assert d(cams["camera_initRec"]["tran"][i], want["camera_initRec_t"][0]) <= TOL, f"{tag}: initRec tran"
    assert d(cams["camera_initRec"]["rot"][i], want["camera_initRec_q"][0]) <= TOL, f"{tag}: initRec rot"
    # matcher: probabilities abs, raw log-scores relative, assignment exact
    wl = want["log_scores_padded"][0]
    gl = lsp[0][i].cpu()
    assert d(gl.exp(), wl.exp()) <= TOL, f"{tag}: exp(log_scores) off by {d(gl.exp(), wl.exp())}"
    rel = float(((gl - wl).abs() / wl.abs().clamp_min(1.0)).max())
    assert rel <= 5 * util.LOG_REL_TOL, f"{tag}: log-scores rel {rel}"
    assert torch.equal(ass["pred_assignment_beforeRef0"][i].cpu(), want["assignment_before"][0]), f"{tag}: assignment"
    if "assignment_after" in want:
        assert torch.equal(ass["pred_assignment"][i].cpu(), want["assignment_after"][0]), f"{tag}: pruned assignment"
    # geo sequences + sig (discrete)
    assert torch.equal(pro["sig_seq"][i].cpu(), want["sig_seq"]), f"{tag}: sig_seq"
    assert d(pro["geo_local"][i], want["geo_local"]) == 0.0, f"{tag}: geo_local (pure gather) must be exact"
    # geo_global is an internal stage boundary: the warp factor k = 1 + t.b/|b|^2 amplifies the (<=1e-4)
    # deviation of the upstream pose (q0,t0) by |t|/d; the kernel itself is checked on identical inputs in
    # test_geo_sequence_kernel_on_oracle_inputs.
    assert d(pro["geo_global"][i], want["geo_global"]) <= 5e-3, f"{tag}: geo_global {d(pro['geo_global'][i], want['geo_global'])}"
    # refined poses
    assert d(cams["camera"]["tran"][i], want["pred_trans"][0]) <= TOL, f"{tag}: camera tran {d(cams['camera']['tran'][i], want['pred_trans'][0])}"
    assert d(cams["camera"]["rot"][i], want["pred_rot"][0]) <= TOL, f"{tag}: camera rot {d(cams['camera']['rot'][i], want['pred_rot'][0])}"
    assert d(cams["camera_avgRef0"]["tran"][i], want["pred_trans_avg"][0]) <= TOL, f"{tag}: avg tran"
    assert d(cams["camera_avgRef0"]["rot"][i], want["pred_rot_avg"][0]) <= TOL, f"{tag}: avg rot"
    if "all_pred_rots" in want:
        assert d(pro["all_pred_rots"][i, :m + 1], want["all_pred_rots"][0]) <= TOL, f"{tag}: all_pred_rots"
        assert d(pro["all_pred_trans"][i, :m + 1], want["all_pred_trans"][0]) <= TOL, f"{tag}: all_pred_trans"
        sr, st = pro["score_soft_rot"][i].cpu(), pro["score_soft_offset"][i].cpu()
        assert d(sr[:m + 1], want["score_soft_rot"][0, :, 0]) <= TOL, f"{tag}: score_soft_rot"
        assert d(st[:m + 1], want["score_soft_offset"][0, :, 0]) <= TOL, f"{tag}: score_soft_offset"
        if m + 1 < sr.numel():
            assert float(sr[m + 1:].abs().max()) == 0.0 and float(st[m + 1:].abs().max()) == 0.0, f"{tag}: padded scores"
    if "l2_dist" in want and "l2_dist" in pro:
        # diagnostics only (sample-0 outputs of the reference); distances are unbounded -> abs + rel bar
        close = lambda a, b, atol: torch.allclose(a.cpu(), b, rtol=1e-4, atol=atol)
        assert close(pro["l2_dist"][i, :m + 1, :m], want["l2_dist"][0], TOL), f"{tag}: l2_dist"
        assert close(pro["normal_dist"][i, :m + 1, :m], want["normal_dist"][0], 2e-2), f"{tag}: normal_dist (deg; acos is ill-conditioned near 0)"
        assert close(pro["offset_dist"][i, :m + 1, :m], want["offset_dist"][0], TOL), f"{tag}: offset_dist"


def _selection_from_reference(want, case):
    """Index the reference picked in min-cost / max-score mode = the row of all_pred_* equal to its output."""
    rows_r = (want["all_pred_rots"][0] == want["pred_rot"][0]).all(-1).nonzero()[:, 0]
    rows_t = (want["all_pred_trans"][0] == want["pred_trans"][0]).all(-1).nonzero()[:, 0]
    return int(rows_r[0]), int(rows_t[0])


@pytest.mark.parametrize("name", util.golden_names())
def test_cuda_matches_golden(name):
    g = util.load_golden(name)
    case = g["case"]
    pairs = case["pairs"]
    cams, _, _, lsp, ass, pro = _run_cuda_case(case, pairs, want_diag=case["NQ"] <= 64)
    for i, (pi, want) in enumerate(zip(pairs, g["outputs"])):
        _check_against(want, cams, lsp, ass, pro, i, f"{name}[pair {pi}]")
        if case["cam"] in ("min-cost", "max-score") and int(want["matched_num"]) > 1:
            sr, st = _selection_from_reference(want, case)
            got = pro["sel_idx"][i].cpu().tolist()
            assert got == [sr, st], f"{name}[pair {pi}]: selection {got} != reference {[sr, st]}"


def test_batched_equals_per_pair_oracle():
    """A batch with DIFFERENT hypothesis counts per pair (0, 1, 5, 16, 48) in one launch equals the per-pair
    loop of the oracle — the [0]-indexed shortcuts of the reference (:964, :1052, :1068) are applied per
    sample (SURVEY.md §7 'Batch semantics' / 'Variable m per pair')."""
    dev = _gpu()
    from oracle import restate
    from nopesac_b200 import synthetic
    NQ, P, B = 50, 16, 6
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    b = synthetic.make_batch(20, B, P)
    over = torch.zeros(B, P, P)
    over[1, 3, 7] = 1
    over[2, [0, 2, 5, 9, 15], [4, 4, 1, 0, 15]] = 1
    over[3] = torch.eye(P)
    over[4, :3] = 1
    over[5, torch.arange(P), b.perm[5]] = 1
    poses = [util.initial_pose_for(100 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    outs = []
    with torch.no_grad():
        for i in range(B):
            outs.append(restate.inference_joint(sd, msd, None, None, b.planes1[i:i + 1], b.planes2[i:i + 1],
                                                b.app1[i:i + 1], b.app2[i:i + 1], num_queries=NQ,
                                                hyp_pairs=torch.nonzero(over[i]), initial_pose=(ip[0][i:i + 1], ip[1][i:i + 1])))
    assert [o["matched_num"] for o in outs] == [0, 1, 5, 16, 48, 16]
    bd = b.to(dev)
    cams, _, _, lsp, ass, pro = head(None, None, bd.planes1, bd.planes2, bd.app1, bd.app2, matching_net=match,
                                     initial_pose=(ip[0].to(dev), ip[1].to(dev)), assignment_override=over.to(dev))
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        _check_against(util.oracle_to_flat(o), cams, lsp, ass, pro, i, f"ragged batch pair {i} (m={o['matched_num']})")


def _oracle_stage_inputs(nq=50, planes=16, pair=4, hyp=None):
    """Stage-boundary tensors of the oracle for one pair (kernel-level tests run on IDENTICAL inputs)."""
    from nopesac_b200 import synthetic
    sd, msd = util.make_weights(nq)
    b = synthetic.make_batch(pair, 1, planes)
    ip = util.initial_pose_for(pair)
    from oracle import restate
    with torch.no_grad():
        o = restate.inference_joint(sd, msd, None, None, b.planes1, b.planes2, b.app1, b.app2, num_queries=nq,
                                    hyp_pairs=hyp, initial_pose=ip)
    return sd, msd, b, o


def test_geo_sequence_kernel_on_oracle_inputs():
    """K6 alone: same (t0, q0, assignment) as the oracle -> geo_local exact, geo_global / 8-vector 1e-5,
    sig_seq / matched_num / nonzero order exact."""
    dev = _gpu()
    from nopesac_b200 import ops
    for negk in (False, True):
        from nopesac_b200 import synthetic
        sd, msd = util.make_weights(50)
        b = synthetic.make_batch(3, 1, 16, negative_k=negk)
        ip = util.initial_pose_for(3)
        from oracle import restate
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, None, None, b.planes1, b.planes2, b.app1, b.app2, num_queries=50,
                                        initial_pose=ip)
        t0, q0 = o["camera_initRec"]
        gl, gg, sig, geo8, mnum, pidx = ops.geo_sequence(b.planes1.to(dev), b.planes2.to(dev),
                                                          o["assignment_before"].to(dev), t0.to(dev), q0.to(dev), 50)
        m = o["matched_num"]
        assert int(mnum[0]) == m
        assert torch.equal(pidx[0, :m].cpu().long(), torch.nonzero(o["assignment_before"][0]))
        assert bool((pidx[0, m:] == -1).all())
        assert torch.equal(gl[0].cpu(), o["geo_local"])
        assert util.maxdiff(gg[0], o["geo_global"]) <= 2e-5
        assert torch.equal(sig[0].cpu(), o["sig_seq"][:, 0])
        if negk:
            assert float(sig[0].min()) == -1.0, "negative-k case must exercise sig_seq = -1"


def test_score_kernel_on_oracle_inputs():
    """K8+K9 alone on the oracle's own per-hypothesis features: scores / poses 1e-5, selections exact,
    for every out_cam_type."""
    dev = _gpu()
    from nopesac_b200 import ops, synthetic
    from oracle import restate
    nq = 64
    hyp = synthetic.all_pairs_hypotheses(16, 40)
    sd, msd, b, o = _oracle_stage_inputs(nq, 16, 1, hyp)
    t0, q0 = o["camera_initRec"]
    with torch.no_grad():
        _, rf0 = restate.rot_rec_head(sd, o["camera_init"][1])
        _, tf0 = restate.trans_rec_head(sd, o["camera_init"][0])
        fr, ft = restate.hypothesis_features(sd, o["geo_global"][None], o["sig_seq"][None], rf0, tf0)
        qh = torch.nn.functional.normalize(restate.linear(sd, "rots", fr), dim=-1)
        th = restate.linear(sd, "trans", ft)
    c = lambda x: x.to(dev).contiguous()
    mlp = lambda p, r: tuple(c(sd[k]) for k in (f"{p}.layers.0.weight", f"{p}.layers.0.bias", f"{p}.layers.1.weight",
                                                f"{p}.layers.1.bias", f"{p}.layers.2.weight", f"{p}.layers.2.bias",
                                                f"{r}.weight", f"{r}.bias"))
    for cam in ("soft", "avg-all", "min-cost", "max-score"):
        with torch.no_grad():
            want = restate.score_and_select(sd, fr, ft, rf0, tf0, o["geo_local"][None], o["matched_num"], q0, t0, cam)
        # exact CUDA-core path (1e-5) and tensor-core path (score MLPs single-pass fp16: 1e-4 bar)
        for precision, tol in (("fp32", 1e-5), ("fp16", 1e-4)):
            res = ops.score_aggregate(c(o["geo_local"][None]), c(qh[None]), c(th[None]), c(q0), c(t0), c(fr[None]), c(ft[None]),
                                      c(rf0), c(tf0), torch.tensor([o["matched_num"]], dtype=torch.int32, device=dev),
                                      mlp("normal_score_proj", "rot_score_reg"), mlp("param_score_proj", "trans_score_reg"),
                                      c(sd["rots.weight"]), c(sd["rots.bias"]), c(sd["trans.weight"]), c(sd["trans.bias"]),
                                      out_cam_type=cam, want_diag=precision == "fp32", precision=precision)
            torch.cuda.synchronize()
            m = o["matched_num"]
            pose = res["pose"][0].cpu()
            tag = f"{cam}/{precision}"
            assert util.maxdiff(pose[0:3], want["pred_trans"][0]) <= tol and util.maxdiff(pose[3:7], want["pred_rot"][0]) <= tol, tag
            assert util.maxdiff(pose[7:10], want["pred_trans_avg"][0]) <= tol and util.maxdiff(pose[10:14], want["pred_rot_avg"][0]) <= tol, tag
            assert int(pose[14]) == m
            assert util.maxdiff(res["score_rot"][0, :m + 1], want["score_soft_rot"][0, :, 0]) <= tol, tag
            assert util.maxdiff(res["score_tran"][0, :m + 1], want["score_soft_offset"][0, :, 0]) <= tol, tag
            assert float(res["score_rot"][0, m + 1:].abs().max()) == 0.0
            if cam in ("min-cost", "max-score"):
                assert res["sel_idx"][0].cpu().tolist() == [want["sel_rot"], want["sel_tran"]], tag
            if precision == "fp32":
                assert torch.allclose(res["diag"][0, 0, :m + 1, :m].cpu(), want["l2_dist"][0], rtol=1e-5, atol=1e-5)


def test_score_tc_kernel_ragged_batch_vs_exact_kernel():
    """The tensor-core scoring path against the exact CUDA-core path on a batch with every m regime
    (0, 1, 2, 63, 64, 65, 128, 129, 255, 256 of NQ = 256) — tiles, k-block tails and the empty cases."""
    dev = _gpu()
    from nopesac_b200 import ops
    NQ = 256
    ms = [0, 1, 2, 63, 64, 65, 128, 129, 255, 256]
    B = len(ms)
    head, _, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    pk = head.prepare()
    g = torch.Generator(device=dev).manual_seed(11)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    geo = rnd(B, NQ, 6)
    qh = torch.nn.functional.normalize(rnd(B, NQ, 4), dim=-1)
    th = rnd(B, NQ, 3) * 0.3
    q0 = torch.nn.functional.normalize(rnd(B, 4), dim=-1)
    t0 = rnd(B, 3) * 0.3
    # feature scale 0.3 keeps the regressed poses O(1), the regime the 1e-4 absolute bar is defined for
    fr, ft, fr0, ft0 = rnd(B, NQ, 256) * 0.3, rnd(B, NQ, 256) * 0.3, rnd(B, 256) * 0.3, rnd(B, 256) * 0.3
    mnum = torch.tensor(ms, device=dev, dtype=torch.int32)
    for i, m in enumerate(ms):       # padded rows are zero in the real pipeline
        geo[i, m:] = 0
    for cam in ("soft", "min-cost", "max-score"):
        out = {}
        for precision in ("fp32", "fp16"):
            out[precision] = ops.score_aggregate(geo, qh, th, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"],
                                                 pk["param_score_proj"], head.rots.weight, head.rots.bias, head.trans.weight,
                                                 head.trans.bias, out_cam_type=cam, precision=precision)
        torch.cuda.synchronize()
        a, b = out["fp32"], out["fp16"]
        assert util.maxdiff(a["pose"], b["pose"]) <= 1e-4, (cam, util.maxdiff(a["pose"], b["pose"]))
        assert util.maxdiff(a["score_rot"], b["score_rot"]) <= 1e-4 and util.maxdiff(a["score_tran"], b["score_tran"]) <= 1e-4, cam
        if cam == "min-cost":      # distances are fp32 on both paths: same argmin
            assert torch.equal(a["sel_idx"], b["sel_idx"]), cam


@pytest.mark.parametrize("B,NQ", [(130, 64), (7, 300), (257, 50)])
def test_score_tc_kernel_row0_tiles_and_odd_shapes(B, NQ):
    """Tensor-core scoring path vs the exact CUDA-core path where the work-item list is irregular: more than one row-0 tile
    with a partial last one (B = 130, 257: hypothesis 0 of 128 pairs per tile), NQ not a multiple of 64 / three hypothesis
    tiles per pair (NQ = 300 -> NQp = 320), random m per pair including 0 and NQ."""
    dev = _gpu()
    from nopesac_b200 import ops
    head, _, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    pk = head.prepare()
    g = torch.Generator(device=dev).manual_seed(100 + B)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    geo = rnd(B, NQ, 6)
    qh = torch.nn.functional.normalize(rnd(B, NQ, 4), dim=-1)
    th = rnd(B, NQ, 3) * 0.3
    q0 = torch.nn.functional.normalize(rnd(B, 4), dim=-1)
    t0 = rnd(B, 3) * 0.3
    fr, ft, fr0, ft0 = rnd(B, NQ, 256) * 0.3, rnd(B, NQ, 256) * 0.3, rnd(B, 256) * 0.3, rnd(B, 256) * 0.3
    mnum = torch.randint(0, NQ + 1, (B,), device=dev, generator=g, dtype=torch.int32)
    mnum[0], mnum[-1] = NQ, 0
    if B > 2:
        mnum[1] = 1
    valid = torch.arange(NQ, device=dev)[None, :] < mnum[:, None]
    geo = geo * valid[:, :, None]            # padded rows are zero in the real pipeline
    for cam in ("soft", "avg-all", "min-cost", "max-score"):
        out = {}
        for precision in ("fp32", "fp16"):
            out[precision] = ops.score_aggregate(geo, qh, th, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"],
                                                 pk["param_score_proj"], head.rots.weight, head.rots.bias, head.trans.weight,
                                                 head.trans.bias, out_cam_type=cam, precision=precision)
        torch.cuda.synchronize()
        a, b = out["fp32"], out["fp16"]
        assert util.maxdiff(a["score_rot"], b["score_rot"]) <= 1e-4 and util.maxdiff(a["score_tran"], b["score_tran"]) <= 1e-4, cam
        assert torch.equal(a["pose"][:, 14], mnum.float()) and torch.equal(b["pose"][:, 14], mnum.float())
        lin = [0, 1, 2, 7, 8, 9, 10, 11, 12, 13]      # translations (linear in the features) and the score-free average pose
        assert util.maxdiff(a["pose"][:, lin], b["pose"][:, lin]) <= 1e-4, (cam, util.maxdiff(a["pose"][:, lin], b["pose"][:, lin]))
        # The soft quaternion is normalize(W_rots . sum_h s_h f_h + b): with RANDOM features the sum nearly cancels, so the
        # normalisation divides by a small norm and amplifies the (<= 2e-5) score differences of the fp16 score MLPs.  The
        # bar is therefore set on the un-normalised vector: |dq| * |u| <= 1e-4 (|u| from the exact path's scores, in torch).
        with torch.no_grad():
            H = NQ + 1
            feats = torch.cat([fr0[:, None], fr], 1)                                   # hypothesis 0 first
            live = torch.arange(H, device=dev)[None, :] <= mnum[:, None]
            agg = ((a["score_rot"] * live)[:, :, None] * feats).sum(1)
            u = agg @ head.rots.weight.T + head.rots.bias
            norm = u.norm(dim=1).clamp(max=1.0)
        dq = (a["pose"][:, 3:7] - b["pose"][:, 3:7]).abs().max(dim=1).values
        assert float((dq * norm).max()) <= 1e-4, (cam, float((dq * norm).max()), float(dq.max()))
        if cam == "min-cost":      # distances are fp32 on both paths: same argmin
            assert torch.equal(a["sel_idx"], b["sel_idx"]), cam


def test_linear_kernel_against_torch():
    """nsac_linear vs torch fp64 on odd shapes (K = 3, 8; N = 3; strided in/out; grouped bias)."""
    dev = _gpu()
    g = torch.Generator().manual_seed(0)
    from nopesac_b200 import ops
    for (M, N, K, act) in ((7, 3, 3, 0), (130, 256, 8, 1), (257, 129, 1024, 2), (64, 512, 1280, 1), (1, 4, 256, 0)):
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        bias = torch.randn(N, generator=g)
        ref = x.double() @ w.double().T + bias.double()
        ref = torch.relu(ref) if act == 1 else (torch.nn.functional.leaky_relu(ref, 0.01) if act == 2 else ref)
        got = ops.linear(x.to(dev), w.to(dev), bias.to(dev), act)
        assert util.maxdiff(got, ref) <= 2e-5, (M, N, K, util.maxdiff(got, ref))
    # strided input / output + grouped bias
    M, N, K, G = 96, 64, 32, 8
    buf = torch.randn(M, 100, generator=g).to(dev)
    w = torch.randn(N, K, generator=g).to(dev)
    gb = torch.randn(M // G, N, generator=g).to(dev)
    out = torch.zeros(M, 200, device=dev)
    ops.linear(buf[:, 4:4 + K], w, gb, 0, out=out[:, 8:8 + N], bias_group_rows=G)
    ref = buf[:, 4:4 + K].double() @ w.double().T + gb.double().repeat_interleave(G, 0)
    assert util.maxdiff(out[:, 8:8 + N], ref) <= 2e-5
    assert float(out[:, :8].abs().max()) == 0.0 and float(out[:, 8 + N:].abs().max()) == 0.0


def test_cabi_error_reporting():
    """Bad arguments return an error code + message instead of crashing (SURVEY.md §8b 'Errors')."""
    _gpu()
    from nopesac_b200 import _lib
    L = _lib.lib()
    st = L.nsac_linear(None, 4, None, None, 0, None, 4, 4, 4, 4, 0, None)
    assert st == -1 and b"null pointer" in L.nsac_last_error()
    with pytest.raises(RuntimeError):
        from nopesac_b200 import ops
        ops.linear(torch.zeros(2, 2), torch.zeros(2, 2))      # CPU tensors: no CPU path

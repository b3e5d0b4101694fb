"""CPU: the SOURCE of the plain-SIMT kernels (csrc/geo.cu, matcher.cu, score.cu, dense.cu, evaluate.cu — no TMA / tcgen05 /
inline PTX) compiled for the host (tests/simt_host: one OS thread per CUDA thread) and driven through the product's own
Python wrappers (`nopesac_b200.ops`, with the library handle, the CUDA-tensor check and the stream getter swapped by the test
fixture) against the oracle.  The build container has no GPU: this is where the kernel logic of the exact-fp32 stages is
executed before a GPU sees it; the same comparisons run on the device in tests/test_gpu_parity.py.  The tensor-core kernels
(gemm_tc.cu, score_tc.cu, pixel.cu's engine calls) cannot run here and stay GPU-only."""
import ctypes as C
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_host"))


from nopesac_b200 import _lib, ops, synthetic  # noqa: E402
from oracle import restate  # noqa: E402
from tests import host_fixture, util  # noqa: E402


@pytest.fixture()
def host_ops(monkeypatch):
    return host_fixture.install(monkeypatch, with_tensor_standins=True)    # stand-ins only for ops.split() in helper code


def test_geo_sequence_source_on_host(host_ops):
    """K6 (camera_head.py:1352-1425, 568-569, 937-957): nonzero order, matched_num, geo_local, sig exact; geo_global 2e-5."""
    for negk in (False, True):
        sd, msd = util.make_weights(50)
        b = synthetic.make_batch(3, 1, 16, negative_k=negk)
        ip = util.initial_pose_for(3)
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, None, None, b.planes1, b.planes2, b.app1, b.app2, num_queries=50, initial_pose=ip)
        t0, q0 = o["camera_initRec"]
        gl, gg, sig, geo8, mnum, pidx = host_ops.geo_sequence(b.planes1, b.planes2, o["assignment_before"], t0, q0, 50)
        m = o["matched_num"]
        assert int(mnum[0]) == m
        assert torch.equal(pidx[0, :m].long(), torch.nonzero(o["assignment_before"][0]))
        assert bool((pidx[0, m:] == -1).all())
        assert torch.equal(gl[0], o["geo_local"])
        assert util.maxdiff(gg[0], o["geo_global"]) <= 2e-5
        assert torch.equal(sig[0], o["sig_seq"][:, 0])
        if negk:
            assert float(sig[0].min()) == -1.0


def test_geo_sequence_explicit_hypothesis_list_on_host(host_ops):
    """H > P^2 / H < P^2 explicit lists (the BASELINE stress shape) + a batch of 3 pairs."""
    P, NQ = 8, 40
    for H in (20, 40):
        hyp = synthetic.all_pairs_hypotheses(P, H)
        b = synthetic.make_batch(11, 3, P)
        t0 = torch.stack([util.initial_pose_for(i)[0][0] for i in range(3)])
        q0 = torch.stack([util.initial_pose_for(i)[1][0] for i in range(3)])
        gl, gg, sig, geo8, mnum, pidx = host_ops.geo_sequence(b.planes1, b.planes2, None, t0, q0, NQ, hyp_pairs=hyp.to(torch.int32))
        for i in range(3):
            want_l, want_g, want_sig = restate.geo_sequences(b.planes1[i], b.planes2[i], hyp.long(), NQ, t0[i:i + 1], q0[i:i + 1])[:3]
            assert int(mnum[i]) == H
            assert torch.equal(gl[i], want_l) and util.maxdiff(gg[i], want_g) <= 2e-5
            assert torch.equal(sig[i], want_sig.reshape(-1))


def test_match_sinkhorn_assign_source_on_host(host_ops):
    """K5 + a8 (matching_head.py:75-128, 228-306; camera_modules.py:15-34): exp(log_scores_padded) 1e-4, assignment exact;
    ragged n1 != n2, an all-dustbin case (threshold 0.999)."""
    g = torch.Generator().manual_seed(5)
    for (n1, n2, thr) in ((16, 16, 0.2), (5, 9, 0.2), (7, 3, 0.999)):
        b = synthetic.make_batch(21, 1, max(n1, n2))
        p1, p2 = b.planes1[:, :n1].contiguous(), b.planes2[:, :n2].contiguous()
        base = torch.randn(1, max(n1, n2), 256, generator=g)
        d1 = (base[:, :n1] * 1.2).contiguous()
        d2 = (base[:, torch.randperm(max(n1, n2), generator=g)[:n2]] * 1.2 + 0.1 * torch.randn(1, n2, 256, generator=g)).contiguous()
        cam = torch.cat([b.gt_tran, b.gt_quat], dim=1)
        bin_score = torch.tensor(1.0)
        off, nrm = restate.match_penalties(p1, p2, cam)
        s = torch.einsum("bnd,bmd->bnm", d1, d2) / 16.0 - off / 4.0 - nrm / 8.0
        want = restate.log_optimal_transport(s, bin_score, 200)
        want_a = restate.get_assignment_matrix(want, thr)
        lsp, assign = host_ops.match_sinkhorn_assign(d1, d2, p1, p2, cam, bin_score, 4.0, 8.0, 200, thr)
        assert util.maxdiff(lsp.exp(), want.exp()) <= 1e-4, (n1, n2)
        assert torch.equal(assign, want_a), (n1, n2)
        if thr > 0.9:
            assert float(assign.sum()) == 0.0


def test_prune_assignment_source_on_host(host_ops):
    b = synthetic.make_batch(31, 2, 16)
    assign = (torch.rand(2, 16, 16, generator=torch.Generator().manual_seed(1)) < 0.3).float()
    pose = torch.zeros(2, 16)
    pose[:, 0:3], pose[:, 3:7] = b.gt_tran, b.gt_quat
    got = host_ops.prune_assignment(assign, b.planes1, b.planes2, pose)
    want = restate.prune_assignment(assign, b.planes1, b.planes2, b.gt_quat, b.gt_tran)
    assert torch.equal(got, want)
    assert 0 < float(got.sum()) < float(assign.sum())


def test_score_aggregate_exact_source_on_host(host_ops):
    """K8 + K9 exact-fp32 path (camera_head.py:964-1115) on the oracle's own per-hypothesis features: poses / scores 1e-5,
    argmin / argmax selections exact, l2 diagnostic, all four INFERENCE_OUT_CAM_TYPEs, and the m = 0 / m = 1 rules."""
    nq = 32
    sd, msd = util.make_weights(nq)
    for n_hyp in (20, 1, 0):
        b = synthetic.make_batch(3, 1, 8)
        ip = util.initial_pose_for(3)
        hyp = synthetic.all_pairs_hypotheses(8, max(n_hyp, 1))
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, None, None, b.planes1, b.planes2, b.app1, b.app2, num_queries=nq, hyp_pairs=hyp, initial_pose=ip)
            t0, q0 = o["camera_initRec"]
            _, rf0 = restate.rot_rec_head(sd, o["camera_init"][1])
            _, tf0 = restate.trans_rec_head(sd, o["camera_init"][0])
            fr, ft = restate.hypothesis_features(sd, o["geo_global"][None], o["sig_seq"][None], rf0, tf0)
            qh = torch.nn.functional.normalize(restate.linear(sd, "rots", fr), dim=-1)
            th = restate.linear(sd, "trans", ft)
        m = o["matched_num"] if n_hyp else 0
        geo_local = o["geo_local"].clone()
        geo_local[m:] = 0
        mlp = lambda p, r: tuple(sd[k].contiguous() for k in (f"{p}.layers.0.weight", f"{p}.layers.0.bias", f"{p}.layers.1.weight",
                                                               f"{p}.layers.1.bias", f"{p}.layers.2.weight", f"{p}.layers.2.bias",
                                                               f"{r}.weight", f"{r}.bias"))
        for cam in ("soft", "avg-all", "min-cost", "max-score"):
            with torch.no_grad():
                want = restate.score_and_select(sd, fr, ft, rf0, tf0, geo_local[None], m, q0, t0, cam)
            res = host_ops.score_aggregate(geo_local[None].contiguous(), qh[None].contiguous(), th[None].contiguous(), q0, t0,
                                           fr[None].contiguous(), ft[None].contiguous(), rf0, tf0, torch.tensor([m], dtype=torch.int32),
                                           mlp("normal_score_proj", "rot_score_reg"), mlp("param_score_proj", "trans_score_reg"),
                                           sd["rots.weight"], sd["rots.bias"], sd["trans.weight"], sd["trans.bias"],
                                           out_cam_type=cam, want_diag=True, precision="fp32")
            pose, tag = res["pose"][0], f"{cam} m={m}"
            assert util.maxdiff(pose[0:3], want["pred_trans"][0]) <= 1e-5 and util.maxdiff(pose[3:7], want["pred_rot"][0]) <= 1e-5, tag
            assert int(pose[14]) == m, tag
            if m > 0:
                assert util.maxdiff(pose[7:10], want["pred_trans_avg"][0]) <= 1e-5 and util.maxdiff(pose[10:14], want["pred_rot_avg"][0]) <= 1e-5, tag
            if m > 1:
                assert util.maxdiff(res["score_rot"][0, :m + 1], want["score_soft_rot"][0, :, 0]) <= 1e-5, tag
                assert util.maxdiff(res["score_tran"][0, :m + 1], want["score_soft_offset"][0, :, 0]) <= 1e-5, tag
                if cam in ("min-cost", "max-score"):
                    assert res["sel_idx"][0].tolist() == [want["sel_rot"], want["sel_tran"]], tag
                assert torch.allclose(res["diag"][0, 0, :m + 1, :m], want["l2_dist"][0], rtol=1e-5, atol=1e-5), tag


def test_dense_kernels_source_on_host(host_ops):
    """nsac_linear (bias / grouped bias / activations / strided views), nsac_layernorm (+ residual), nsac_attention, nsac_pose_heads."""
    g = torch.Generator().manual_seed(2)
    for (M, N, K, act) in ((5, 7, 3, 0), (33, 256, 8, 1), (64, 65, 130, 2), (1, 4, 256, 0)):
        x, w, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
        want = torch.nn.functional.linear(x, w, bias)
        want = {0: want, 1: want.relu(), 2: torch.nn.functional.leaky_relu(want, 0.01)}[act]
        assert util.maxdiff(host_ops.linear(x, w, bias, act), want) <= 2e-5 * max(1.0, float(want.abs().max())), (M, N, K, act)
    x, res_in = torch.randn(10, 256, generator=g), torch.randn(10, 256, generator=g)
    gamma, beta = torch.randn(256, generator=g), torch.randn(256, generator=g)
    want = res_in + torch.nn.functional.layer_norm(x, (256,), gamma, beta)
    assert util.maxdiff(host_ops.layernorm(x, gamma, beta, res=res_in), want) <= 1e-5
    B, Lq, S, H, D = 2, 5, 9, 8, 32
    q, k, v = (torch.randn(B, n, H * D, generator=g) for n in (Lq, S, S))
    A = torch.softmax(torch.einsum("nlhd,nshd->nlsh", q.view(B, Lq, H, D), k.view(B, S, H, D)) / math.sqrt(D), dim=2)
    want = torch.einsum("nlsh,nshd->nlhd", A, v.view(B, S, H, D)).reshape(B, Lq, H * D)
    got = host_ops.attention(q.view(B * Lq, H * D), k.view(B * S, H * D), v.view(B * S, H * D), B, Lq, S, H, D)
    assert util.maxdiff(got, want.reshape(B * Lq, H * D)) <= 1e-5
    fr, ft = torch.randn(6, 256, generator=g), torch.randn(6, 256, generator=g)
    wr, br, wt, bt = torch.randn(4, 256, generator=g) * 0.1, torch.randn(4, generator=g), torch.randn(3, 256, generator=g) * 0.1, torch.randn(3, generator=g)
    qo, to = host_ops.pose_heads(fr, ft, wr, br, wt, bt)
    assert util.maxdiff(qo, torch.nn.functional.normalize(torch.nn.functional.linear(fr, wr, br), dim=-1)) <= 1e-5
    assert util.maxdiff(to, torch.nn.functional.linear(ft, wt, bt)) <= 1e-5


def test_camera_errors_source_on_host(host_ops):
    """Row f3 kernel (mp3d_evaluation.py:382-425, 463-465) against the evaluation oracle and the golden fixture."""
    import json
    import numpy as np
    from oracle import eval_restate
    L = _lib._lib
    with open(os.path.join(ROOT, "tests", "golden", "camera_eval.json")) as f:
        cases = json.load(f)
    for c in cases:
        n = c["n"]
        arr = lambda k: np.ascontiguousarray(np.asarray(c[k], dtype=np.float32))
        rows = np.zeros((n, 16), np.float32)
        rows[:, 0:3], rows[:, 3:7] = arr("pred_tran"), arr("pred_rot")
        gt_t, gt_q = arr("gt_tran"), arr("gt_rot")
        et, er, stats = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(8, np.float32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        assert L.nsac_camera_errors(p(rows), 16, p(gt_t), p(gt_q), n, p(et), p(er), p(stats), None) == 0
        assert np.abs(et - np.linalg.norm(gt_t - rows[:, 0:3], axis=1)).max() <= 1e-5
        assert np.abs(er - eval_restate.angle_error_vec(rows[:, 3:7], gt_q)).max() <= 5e-2
        want = c["metrics"]
        assert abs(stats[0] - want["T mean err"]) <= 1e-5 and abs(stats[1] - want["R mean err"]) <= 2e-3
        for i, k in enumerate(("T err < 1.0", "T err < 0.5", "T err < 0.2", "R err < 30", "R err < 15", "R err < 10")):
            assert abs(float(stats[2 + i]) / n * 100 - want[k]) <= 1e-9, (n, k)


def test_ragged_attention_source_on_host(host_ops):
    """nsac_attention_ragged: batch element b attends to its first kv_count[b] source tokens only == attention on the slice."""
    g = torch.Generator().manual_seed(4)
    B, Lq, S, H, D = 3, 6, 9, 8, 32
    q, k, v = (torch.randn(B, n, H * D, generator=g) for n in (Lq, S, S))
    cnt = torch.tensor([9, 4, 1], dtype=torch.int32)
    k[1, 4:], v[1, 4:], k[2, 1:], v[2, 1:] = 1e30, float("nan"), -1e30, float("nan")      # padding must never be touched
    got = host_ops.attention(q.view(B * Lq, H * D), k.view(B * S, H * D), v.view(B * S, H * D), B, Lq, S, H, D, kv_count=cnt).view(B, Lq, H * D)
    for b in range(B):
        n = int(cnt[b])
        A = torch.softmax(torch.einsum("lhd,shd->lsh", q[b].view(Lq, H, D), k[b, :n].view(n, H, D)) / math.sqrt(D), dim=1)
        want = torch.einsum("lsh,shd->lhd", A, v[b, :n].view(n, H, D)).reshape(Lq, H * D)
        assert util.maxdiff(got[b], want) <= 1e-5, b
    full = host_ops.attention(q.view(B * Lq, H * D)[:Lq], k.view(B * S, H * D)[:S], v.view(B * S, H * D)[:S], 1, Lq, S, H, D)
    assert torch.equal(full, got[0])                       # count == S is the unmasked kernel, bit for bit


def test_ragged_match_sinkhorn_assign_source_on_host(host_ops):
    """nsac_match_sinkhorn_assign_ragged: every pair of a padded batch == the un-padded single-pair call (bit for bit) and the
    oracle on the slices; padding of the outputs is -inf / 0; NaN padding of the inputs is never read."""
    g = torch.Generator().manual_seed(9)
    P = 12
    counts = [(12, 12), (5, 9), (7, 3), (1, 4)]
    B = len(counts)
    b = synthetic.make_batch(40, B, P)
    base = torch.randn(B, P, 256, generator=g)
    d1 = base * 1.2
    d2 = torch.stack([base[i, torch.randperm(P, generator=g)] for i in range(B)]) * 1.2 + 0.1 * torch.randn(B, P, 256, generator=g)
    p1, p2 = b.planes1.clone(), b.planes2.clone()
    for i, (a, c) in enumerate(counts):
        d1[i, a:], p1[i, a:], d2[i, c:], p2[i, c:] = float("nan"), float("nan"), float("nan"), float("nan")
    cam = torch.cat([b.gt_tran, b.gt_quat], dim=1)
    bin_score = torch.tensor(1.0)
    c1 = torch.tensor([a for a, _ in counts], dtype=torch.int32)
    c2 = torch.tensor([c for _, c in counts], dtype=torch.int32)
    lsp, assign = host_ops.match_sinkhorn_assign(d1, d2, p1, p2, cam, bin_score, 4.0, 8.0, 200, 0.2, count1=c1, count2=c2)
    for i, (a, c) in enumerate(counts):
        s1, s2 = (d1[i:i + 1, :a].contiguous(), d2[i:i + 1, :c].contiguous())
        q1, q2 = p1[i:i + 1, :a].contiguous(), p2[i:i + 1, :c].contiguous()
        one_lsp, one_assign = host_ops.match_sinkhorn_assign(s1, s2, q1, q2, cam[i:i + 1].contiguous(), bin_score, 4.0, 8.0, 200, 0.2)
        assert torch.equal(lsp[i, :a + 1, :c + 1], one_lsp[0]) and torch.equal(assign[i, :a, :c], one_assign[0]), i
        assert bool(torch.isinf(lsp[i, a + 1:]).all()) and bool(torch.isinf(lsp[i, :, c + 1:]).all())
        assert float(assign[i, a:].abs().sum()) == 0.0 and float(assign[i, :, c:].abs().sum()) == 0.0
        off, nrm = restate.match_penalties(q1, q2, cam[i:i + 1])
        s = torch.einsum("bnd,bmd->bnm", s1, s2) / 16.0 - off / 4.0 - nrm / 8.0
        want = restate.log_optimal_transport(s, bin_score, 200)
        assert util.maxdiff(one_lsp.exp(), want.exp()) <= 1e-4, i
        assert torch.equal(one_assign, restate.get_assignment_matrix(want, 0.2)), i


def test_entry_points_reject_bad_arguments_with_a_message(host_ops):
    """SURVEY.md §8b 'Errors': every export returns a negative status and sets nsac_last_error() instead of crashing — checked
    on the host build of the same sources (argument validation happens before any launch)."""
    L = _lib._lib
    L.nsac_last_error.restype = C.c_char_p
    err = lambda: L.nsac_last_error().decode()
    z = torch.zeros(64)
    p = lambda t: C.c_void_p(t.data_ptr())
    assert L.nsac_linear(None, 4, None, None, 0, None, 4, 4, 4, 4, 0, None) == -1 and "null pointer" in err()
    assert L.nsac_attention(p(z), 256, p(z), p(z), 256, p(z), 256, None, None, 0, 1, 1, 1, 8, 16, None) == -1 and "head dim must be 32" in err()
    assert L.nsac_match_sinkhorn_assign(p(z), p(z), p(z), p(z), p(z), p(z), 4.0, 8.0, 200, 0.2, 1, 0, 3, 256, p(z), p(z), None) == -1
    assert "bad shape" in err()
    assert L.nsac_match_sinkhorn_assign(p(z), p(z), p(z), p(z), p(z), p(z), 4.0, 8.0, 200, 0.2, 1, 2000, 3, 256, p(z), p(z), None) == -1
    assert "1023 planes" in err()
    assert L.nsac_prune_assignment(p(z), p(z), p(z), p(z), 3, 1, 2, 2, p(z), None) == -1 and "bad shape" in err()
    assert L.nsac_camera_errors(p(z), 3, p(z), p(z), 1, p(z), p(z), p(z), None) == -1 and "bad shape" in err()
    assert L.nsac_geo_sequence(None, None, None, None, 0, None, None, 1, 2, 2, 8, None, None, None, None, None, None, None) == -1
    assert L.nsac_plane_post_workspace_bytes(1, 0, 480, 640) == 0
    # a successful call afterwards is unaffected by the stale message
    x, w = torch.randn(3, 4), torch.randn(5, 4)
    assert util.maxdiff(host_ops.linear(x, w), x @ w.T) <= 1e-6


def test_backbone_glue_kernels_source_on_host(host_ops):
    """csrc/backbone.cu (row f2): stem normalisation + 7x7/2 im2col, MaxPool2d(3,2,1), stride-2 subsampling, relu(a + b) against
    PyTorch; odd and even sizes."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(6)
    nhwc = lambda x: x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    for (N, H, W) in ((2, 16, 24), (1, 15, 21)):
        img = torch.rand(N, 3, H, W, generator=g) * 255
        cols, Ho, Wo = host_ops.stem_im2col_planes(img, mean, std)
        norm = (img - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
        ref = F.unfold(norm, 7, padding=3, stride=2)                                  # [N, 3*49, L], rows (c, ky, kx)
        assert (Ho, Wo) == ((H - 1) // 2 + 1, (W - 1) // 2 + 1) and ref.shape[2] == Ho * Wo
        ref = ref.reshape(N, 3, 49, Ho * Wo).permute(0, 3, 2, 1).reshape(N * Ho * Wo, 147)      # (ky,kx,c) order
        assert util.maxdiff(cols.float(), ref) <= 2e-6
        assert float((cols.hi[:, 147:].float().abs() + cols.lo[:, 147:].float().abs()).max()) == 0.0
        w = torch.randn(64, 3, 7, 7, generator=g)
        conv = F.conv2d(norm.double(), w.double(), stride=2, padding=3)
        assert util.maxdiff(cols.float().double() @ w.permute(0, 2, 3, 1).reshape(64, 147).double().T, nhwc(conv)) <= 1e-4
        x = torch.randn(N, 64, H, W, generator=g)
        out, sp, Hp, Wp = host_ops.maxpool3x3s2_nhwc(nhwc(x), N, H, W)
        ref = F.max_pool2d(x, 3, 2, 1)
        assert ref.shape[2:] == (Hp, Wp) and torch.equal(out, nhwc(ref)) and util.maxdiff(sp.float(), out) <= 1e-6
        xp = host_ops.split(nhwc(x))
        sub, Hs, Ws = host_ops.subsample2_planes(xp, N, H, W)
        assert (Hs, Ws) == ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
        assert util.maxdiff(sub.float(), nhwc(x[:, :, ::2, ::2])) <= 1e-6
        a, b = nhwc(x), nhwc(torch.randn(N, 64, H, W, generator=g))
        o, op_ = host_ops.add_relu_nhwc(a, b)
        assert torch.equal(o, torch.relu(a + b)) and util.maxdiff(op_.float(), o) <= 1e-6

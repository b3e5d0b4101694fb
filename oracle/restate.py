"""TEST INFRASTRUCTURE — the parity oracle.  Not product code; never imported by nopesac_b200/.

CPU (plain PyTorch, eager) restatement of the reference's one-plane RANSAC pose path, op for op and
in the reference's operation order, so that (a) it is the checker for the CUDA path in tests/,
smoke() and (b) its wall-clock on the host cores is the `cpu_baseline` ("kind": "port") of bench.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Pinning: the reference ships NO golden vectors or tests, so this restatement is pinned
against the *live reference code* imported from /root/reference (oracle/ref_loader.py) in
tests/test_oracle_vs_reference.py (runs wherever /root/reference exists) and against fixtures generated
from that live reference and committed under tests/golden/ (tests/golden/make_golden.py).

Each function cites the reference lines it follows (paths relative to
/root/reference/NopeSAC_Net/modeling/).  The reference supports batch size 1 only; everything here
works on ONE pair (leading dim 1) and `inference_joint_batch` loops pairs — that per-pair loop is the
definition of batched behaviour the CUDA path is held to (SURVEY.md §7 "Batch semantics").
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
from torch.nn import functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------
def mlp(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """camera_net/camera_modules.py:226-244 — Linear+ReLU, last layer linear."""
    n = 0
    while f"{prefix}.layers.{n}.weight" in sd:
        n += 1
    for i in range(n):
        x = F.linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def quaternion2rotmatrix(q: torch.Tensor) -> torch.Tensor:
    """camera_net/camera_head.py:1135-1177 ([n,4] (w,x,y,z) -> [n,3,3])."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros(q.shape[0], 3, 3, dtype=q.dtype)
    R[:, 0, 0] = 1 - 2 * y * y - 2 * z * z
    R[:, 0, 1] = 2 * x * y - 2 * w * z
    R[:, 0, 2] = 2 * x * z + 2 * w * y
    R[:, 1, 0] = 2 * x * y + 2 * w * z
    R[:, 1, 1] = 1 - 2 * x * x - 2 * z * z
    R[:, 1, 2] = 2 * y * z - 2 * w * x
    R[:, 2, 0] = 2 * x * z - 2 * w * y
    R[:, 2, 1] = 2 * y * z + 2 * w * x
    R[:, 2, 2] = 1 - 2 * x * x - 2 * y * y
    return R


def _flip(dtype):
    return torch.tensor([1.0, -1.0, -1.0], dtype=dtype).reshape(1, 1, 3)


def warp_to_global(plane: torch.Tensor, rot_quan=None, tran=None, pose_n=None) -> torch.Tensor:
    """camera_net/camera_head.py:1427-1466 (twin: matching_net/matching_head.py:141-180).
    plane [bs,n,3]; rot_quan [bs,h,4]; tran [bs,h,3] -> [bs,h,n,3]; or the view-2 flip when no pose."""
    if rot_quan is not None and tran is not None:
        bs, h, _ = rot_quan.shape
        n = plane.shape[1]
        plane0 = plane.unsqueeze(1).repeat(1, h, 1, 1).view(bs * h, n, 3)
        R = quaternion2rotmatrix(rot_quan.reshape(-1, 4))
        tr = tran.unsqueeze(2).repeat(1, 1, n, 1).view(bs * h, n, 3)
        end = (plane0 * _flip(plane.dtype)).permute(0, 2, 1)
        end = torch.bmm(R, end).permute(0, 2, 1) + tr
        b = end - tr
        out = ((end * b).sum(dim=-1) / (torch.norm(b, dim=-1) + 1e-5) ** 2).view(bs * h, n, 1) * b
        return out.reshape(bs, h, n, 3).contiguous()
    bs, n = plane.shape[:2]
    plane1 = plane.unsqueeze(1).repeat(1, pose_n, 1, 1).view(bs * pose_n, n, 3) * _flip(plane.dtype)
    return plane1.reshape(bs, pose_n, n, 3).contiguous()


# ----------------------------------------------------------------------------------------------
# K1 — pixel pose regression network (camera_head.py:642-683)
# ----------------------------------------------------------------------------------------------
def _conv_block(sd: SD, prefix: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """camera_modules.py:36-48 — Conv3x3(no bias) + BatchNorm2d(eps=1e-3, eval) + LeakyReLU(0.01)."""
    x = F.conv2d(x, sd[prefix + ".0.weight"], None, stride=stride, padding=1)
    x = F.batch_norm(x, sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"],
                     sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], False, 0.0, 1e-3)
    return F.leaky_relu(x, 0.01)


def _d2_conv(sd: SD, prefix: str, x, padding, relu):
    """detectron2.layers.Conv2d: conv (+bias if present) -> GroupNorm(32) if present -> activation."""
    x = F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), padding=padding)
    if prefix + ".norm.weight" in sd:
        x = F.group_norm(x, 32, sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"], 1e-5)
    return F.relu(x) if relu else x


def pixel_decoder_features(sd: SD, feats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """camera_modules.py:335-348 with the module table of :258-322 (res2 dropped; top-down res5->res3)."""
    p = "pixel_decoder."
    y = _d2_conv(sd, p + "layer_3", feats["res5"], 1, True)
    for lvl, name in ((2, "res4"), (1, "res3")):
        cur = _d2_conv(sd, p + f"adapter_{lvl}", feats[name], 0, False)
        y = cur + F.interpolate(y, size=cur.shape[-2:], mode="nearest")
        y = _d2_conv(sd, p + f"layer_{lvl}", y, 1, True)
    return _d2_conv(sd, p + "mask_features", y, 1, False)


def convs_backbone(sd: SD, x):
    """camera_head.py:78-91."""
    x = _conv_block(sd, "convs_backbone.0", x)
    x = _conv_block(sd, "convs_backbone.1", x)
    x = F.max_pool2d(x, 2, 2)
    x = _conv_block(sd, "convs_backbone.3", x)
    x = _conv_block(sd, "convs_backbone.4", x)
    x = F.max_pool2d(x, 2, 2)
    x = _conv_block(sd, "convs_backbone.6", x)
    x = _conv_block(sd, "convs_backbone.7", x)
    return x


def compute_corr_softmax(f1, f2):
    """camera_head.py:1117-1133."""
    _, _, h1, w1 = f1.shape
    _, _, h2, w2 = f2.shape
    f2v = f2.transpose(2, 3).contiguous().view(f2.size(0), f2.size(1), -1).transpose(1, 2)
    f1v = f1.contiguous().view(f1.size(0), f1.size(1), -1)
    corr = torch.matmul(f2v, f1v).view(f1.size(0), h2 * w2, h1, w1)
    return F.softmax(corr, dim=1)


def pixel_camera_head(sd: SD, feats1, feats2):
    """camera_head.py:642-670 -> (trans [b,3], rot [b,4] normalised, trans_feat, rots_feat)."""
    x1 = convs_backbone(sd, pixel_decoder_features(sd, feats1))
    x2 = convs_backbone(sd, pixel_decoder_features(sd, feats2))
    aff = compute_corr_softmax(x1, x2)
    outs = []
    for br in ("trans", "rots"):
        x = aff
        for i in range(6):
            x = _conv_block(sd, f"convs_{br}.{i}", x, stride=2 if i % 2 == 1 else 1)
        x = F.relu(linear(sd, f"fc_{br}", torch.flatten(x, 1)))
        outs.append(x)
    trans_feat, rots_feat = outs
    trans = linear(sd, "trans", trans_feat)
    rots = F.normalize(linear(sd, "rots", rots_feat), p=2, dim=1)
    return trans, rots, trans_feat, rots_feat


# ----------------------------------------------------------------------------------------------
# K2 — AIM re-embedding (camera_head.py:685-735)
# ----------------------------------------------------------------------------------------------
def rot_rec_head(sd: SD, rot):
    sig = ((rot[:, 0:1] >= 0.).to(rot.dtype) - 0.5) * 2.
    rot = rot * sig
    feat = F.relu(mlp(sd, "rot_emb_proj", rot))
    return F.normalize(linear(sd, "rots", feat), p=2, dim=1), feat


def trans_rec_head(sd: SD, tran):
    tran = tran + 1e-10
    feat = F.relu(mlp(sd, "trans_emb_proj", tran))
    return linear(sd, "trans", feat), feat


# ----------------------------------------------------------------------------------------------
# K3/K4/K5 — matching head (matching_net/matching_head.py, transformer/gnn.py)
# ----------------------------------------------------------------------------------------------
def match_penalties(params1, params2, cam):
    """matching_head.py:75-96: normal angle (deg) and clamped offset distance, [bs,n1,n2] each."""
    p2w = warp_to_global(params2, pose_n=1)[:, 0]
    off2 = torch.norm(p2w, dim=2, keepdim=True, p=2)
    n2 = F.normalize(p2w, dim=-1, p=2)
    q = cam[:, 3:].unsqueeze(1)
    t = cam[:, :3].unsqueeze(1)
    p1r = warp_to_global(params1, q, t * 0.)[:, 0]
    n1r = F.normalize(p1r, dim=-1, p=2)
    nTn_r = torch.bmm(n1r, n2.transpose(1, 2))
    normal_dist = torch.acos(torch.clamp(nTn_r, -1, 1)) / math.pi * 180.
    p1rt = warp_to_global(params1, q, t)[:, 0]
    off1 = torch.norm(p1rt, dim=2, keepdim=True, p=2)
    n1rt = F.normalize(p1rt, dim=-1, p=2)
    nTn_rt = torch.bmm(n1rt, n2.transpose(1, 2))
    offset_dist = torch.abs(off1 - off2.transpose(1, 2))
    neg = nTn_rt < 0
    offset_dist[neg] = torch.abs(off1 + off2.transpose(1, 2))[neg]
    offset_dist = torch.clamp(offset_dist, min=1e-10, max=5.)
    return offset_dist, normal_dist


def gnn_layer(sd: SD, p: str, x, source, nhead=8):
    """transformer/gnn.py:73-96 (+ FullAttention :19-44), no masks at inference."""
    bs, dim = x.size(0), x.size(2) // nhead
    q = F.linear(x, sd[p + "q_proj.weight"]).view(bs, -1, nhead, dim)
    k = F.linear(source, sd[p + "k_proj.weight"]).view(bs, -1, nhead, dim)
    v = F.linear(source, sd[p + "v_proj.weight"]).view(bs, -1, nhead, dim)
    QK = torch.einsum("nlhd,nshd->nlsh", q, k)
    A = torch.softmax((1. / dim ** .5) * QK, dim=2)
    msg = torch.einsum("nlsh,nshd->nlhd", A, v).contiguous()
    msg = F.linear(msg.view(bs, -1, nhead * dim), sd[p + "merge.weight"])
    msg = F.layer_norm(msg, (nhead * dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    msg = F.linear(F.relu(F.linear(torch.cat([x, msg], dim=2), sd[p + "mlp.0.weight"])), sd[p + "mlp.2.weight"])
    msg = F.layer_norm(msg, (nhead * dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return x + msg


def gnn(sd: SD, f0, f1, num_layers=18):
    """transformer/gnn.py:117-138 — ['self','cross']*9; cross(1<-0) sees the UPDATED feat0."""
    for i in range(num_layers):
        p = f"gnn.layers.{i}."
        if i % 2 == 0:
            f0 = gnn_layer(sd, p, f0, f0)
            f1 = gnn_layer(sd, p, f1, f1)
        else:
            f0 = gnn_layer(sd, p, f0, f1)
            f1 = gnn_layer(sd, p, f1, f0)
    return f0, f1


def log_optimal_transport(scores, alpha, iters=200):
    """matching_head.py:259-306 with all-valid masks + :228-234."""
    bsz, M, N = scores.shape
    pc = alpha.expand(bsz, M, 1)
    pr = alpha.expand(bsz, 1, N + 1)
    Z = torch.cat([torch.cat([scores, pc], dim=-1), pr], dim=1)
    nvr = torch.full((bsz,), float(M), dtype=scores.dtype)
    nvc = torch.full((bsz,), float(N), dtype=scores.dtype)
    norm = -torch.log(nvr + nvc)
    log_mu = torch.empty(bsz, M + 1, dtype=scores.dtype)
    log_mu[:, :M] = norm.unsqueeze(1)
    log_mu[:, M] = torch.log(nvc) + norm
    log_nu = torch.empty(bsz, N + 1, dtype=scores.dtype)
    log_nu[:, :N] = norm.unsqueeze(1)
    log_nu[:, N] = torch.log(nvr) + norm
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    out = Z + u.unsqueeze(2) + v.unsqueeze(1)
    return out - norm.unsqueeze(1).unsqueeze(2)


def matching_scores(msd: SD, app1, app2, off_dist, nrm_dist, offset_multiplier=4., normal_multiplier=8.):
    """matching_head.py:101-119: projections, GNN, descriptor similarity minus geometry penalties."""
    a1 = F.conv1d(app1.permute(0, 2, 1), msd["planeApp_proj.weight"], msd["planeApp_proj.bias"]).permute(0, 2, 1)
    a2 = F.conv1d(app2.permute(0, 2, 1), msd["planeApp_proj.weight"], msd["planeApp_proj.bias"]).permute(0, 2, 1)
    d1, d2 = gnn(msd, a1, a2)
    d1 = F.conv1d(d1.permute(0, 2, 1), msd["planeDesc_proj.weight"], msd["planeDesc_proj.bias"])
    d2 = F.conv1d(d2.permute(0, 2, 1), msd["planeDesc_proj.weight"], msd["planeDesc_proj.bias"])
    s = torch.einsum('bdn,bdm->bnm', d1, d2) / 256 ** .5
    s = s - off_dist / offset_multiplier
    s = s - nrm_dist / normal_multiplier
    return s


def matching_head(msd: SD, app1, app2, cam, params1, params2, iters=200):
    """matching_head.py:43-133 (inference branch) -> log_scores_padded [bs,n1+1,n2+1]."""
    off, nrm = match_penalties(params1, params2, cam)
    s = matching_scores(msd, app1, app2, off, nrm)
    return log_optimal_transport(s, msd["bin_score"], iters)


def get_assignment_matrix(log_scores_padded, match_threshold):
    """camera_modules.py:15-34 (bs == 1)."""
    S = log_scores_padded[:, :-1, :-1]
    max0, max1 = S.max(2), S.max(1)
    i0, i1 = max0.indices, max1.indices
    ar0 = torch.arange(i0.shape[1])[None]
    mutual0 = ar0 == i1.gather(1, i0)
    ms0 = torch.where(mutual0, max0.values.exp(), S.new_tensor(0))
    valid0 = mutual0 & (ms0 > match_threshold)
    A = torch.zeros_like(S)
    rows = torch.nonzero(valid0[0])[:, 0]
    A[0, rows, i0[0, rows]] = 1
    return A


# ----------------------------------------------------------------------------------------------
# K6 — geo sequences + sig (camera_head.py:1352-1425, 568-569)
# ----------------------------------------------------------------------------------------------
def geo_sequence_from_pairs(planes1, planes2, idx_pairs, num_queries, cam_t=None, cam_q=None):
    """One pair. planes* [P,3]; idx_pairs [m,2] in torch.nonzero order. camera_head.py:1366-1408."""
    idx1, idx2 = idx_pairs[:, 0], idx_pairs[:, 1]
    m = idx_pairs.shape[0]
    mp1, mp2 = planes1[idx1], planes2[idx2]
    flip = torch.tensor([1., -1., -1.], dtype=planes1.dtype)
    if cam_t is not None:
        R = quaternion2rotmatrix(cam_q.reshape(1, 4)).squeeze(0)
        start = torch.ones((m, 3), dtype=planes1.dtype) * cam_t
        end = torch.mm(R, (mp1 * flip).T).T + cam_t
        b = end - start
        mp1 = ((end * b).sum(dim=1) / (torch.norm(b, dim=1) + 1e-5) ** 2).view(-1, 1) * b
        mp2 = mp2 * flip
    out = torch.zeros(num_queries, 6, dtype=planes1.dtype)
    out[:m] = torch.cat((mp1, mp2), dim=-1)
    return out, m


def geo_sequences(planes1, planes2, idx_pairs, num_queries, t0, q0):
    """camera_head.py:513-517, 555-569 for one pair -> local, global, sig [NQ,1], m."""
    local, m = geo_sequence_from_pairs(planes1, planes2, idx_pairs, num_queries)
    glob, _ = geo_sequence_from_pairs(planes1, planes2, idx_pairs, num_queries, t0, q0)
    aux, _ = geo_sequence_from_pairs(planes1, planes2, idx_pairs, num_queries, torch.zeros_like(t0), q0)
    sig = ((glob[:, 0:1] * aux[:, 0:1]) >= 0).to(planes1.dtype)
    sig = (sig - 0.5) * 2.
    return local, glob, sig, m


# ----------------------------------------------------------------------------------------------
# K7 — hypothesis generation (camera_head.py:934-962, 979-991, 1018-1019)
# ----------------------------------------------------------------------------------------------
def hypothesis_features(sd: SD, geo_global, sig_seq, rot_feat0, trans_feat0):
    """geo_global [bs,n,6], sig [bs,n,1], feats [bs,256] -> fused_rot, fused_trans [bs*n,256]."""
    bs, n, _ = geo_global.shape
    g0 = geo_global[:, :, :3]
    off0 = torch.norm(g0, p=2, dim=-1, keepdim=True)
    nrm0 = g0 / (off0 + 1e-10)
    g1 = geo_global[:, :, 3:]
    off1 = torch.norm(g1, p=2, dim=-1, keepdim=True)
    nrm1 = g1 / (off1 + 1e-10)
    off0 = off0 * sig_seq
    nrm0 = nrm0 * sig_seq
    x = torch.cat((nrm0, off0, nrm1, off1), dim=-1)
    fea = mlp(sd, "geo_encoder", x)
    s1 = mlp(sd, "geo_proj_s1", fea)
    frot = mlp(sd, "decoder_rot", s1)
    s2 = mlp(sd, "geo_proj_s2", torch.cat([s1, frot], dim=-1))
    ftran = mlp(sd, "decoder_tran", s2)
    trans_pad = trans_feat0.unsqueeze(1).repeat(1, n, 1).view(bs * n, -1)
    rot_pad = rot_feat0.unsqueeze(1).repeat(1, n, 1).view(bs * n, -1)
    fused_rot = F.relu(mlp(sd, "decoder_rot2", torch.cat((rot_pad, frot.view(bs * n, -1)), dim=-1)))
    fused_tran = F.relu(mlp(sd, "decoder_tran2", torch.cat((trans_pad, ftran.view(bs * n, -1)), dim=-1)))
    return fused_rot, fused_tran


# ----------------------------------------------------------------------------------------------
# K8/K9 — scoring + selection (camera_head.py:964-1115), one pair (bs == 1)
# ----------------------------------------------------------------------------------------------
def score_and_select(sd: SD, fused_rot, fused_tran, rot_feat0, trans_feat0, geo_local, m, q0, t0,
                     out_cam_type="soft"):
    """fused_* [n,256] (n = NQ), *_feat0 [1,256], geo_local [1,n,6], q0 [1,4], t0 [1,3]."""
    bs, n = 1, geo_local.shape[1]
    dt = geo_local.dtype
    out = {"matched_num": m}
    if m == 0:   # :964-969
        out.update(pred_trans=t0, pred_rot=q0, pred_trans_avg=t0, pred_rot_avg=q0)
        return out
    mask = torch.zeros(bs, n + 1, n, dtype=dt)
    mask[0, :m + 1, :m] = 1.
    # --- rot hypotheses + scoring :990-1014
    rots_all = F.normalize(linear(sd, "rots", fused_rot), dim=-1, p=2).view(bs, n, 4)
    rots_all = torch.cat([q0.unsqueeze(1), rots_all], dim=1)
    zeros_t = torch.zeros(bs, n + 1, 3, dtype=dt)
    p1 = warp_to_global(geo_local[:, :, 3:], pose_n=1).repeat(1, n + 1, 1, 1)
    p0r = warp_to_global(geo_local[:, :, :3], rots_all, zeros_t)
    n0r = F.normalize(p0r, p=2, dim=-1)
    n1r = F.normalize(p1, p=2, dim=-1)
    ang = torch.acos(torch.clamp(torch.sum(n0r * n1r, dim=-1), min=-1., max=1.)) / math.pi * 180.
    dL2 = torch.norm(n0r - n1r, p=2, dim=-1) * mask
    dL2_sum = dL2.sum(-1)
    xr = torch.exp(-dL2) * mask
    sr_tmp = linear(sd, "rot_score_reg", mlp(sd, "normal_score_proj", xr))
    score_rot = torch.zeros_like(sr_tmp)
    score_rot[0, :m + 1] = sr_tmp[0, :m + 1].softmax(0)
    score_rot = score_rot * mask[:, :, 0:1]
    # --- trans hypotheses + scoring :1018-1043
    trans_all = linear(sd, "trans", fused_tran).view(bs, n, 3)
    trans_all = torch.cat([t0.unsqueeze(1), trans_all], dim=1)
    p0rt = warp_to_global(geo_local[:, :, :3], rots_all, trans_all)
    off0 = torch.norm(p0rt, p=2, dim=-1)
    off1 = torch.norm(p1, p=2, dim=-1)
    nTn = torch.sum(F.normalize(p0rt, p=2, dim=-1) * F.normalize(p1, p=2, dim=-1), dim=-1)
    doff = torch.abs(off0 - off1)
    doff[nTn < 0] = torch.abs(off0 + off1)[nTn < 0]
    dl2_ori = torch.norm(p0rt - p1, p=2, dim=-1)
    dl2_sum = (dl2_ori * mask).sum(-1)
    xt = torch.exp(-(dl2_ori * mask)) * mask
    st_tmp = linear(sd, "trans_score_reg", mlp(sd, "param_score_proj", xt))
    score_tran = torch.zeros_like(st_tmp)
    score_tran[0, :m + 1] = st_tmp[0, :m + 1].softmax(0)
    score_tran = score_tran * mask[:, :, 0:1]
    # --- avg :1047-1066
    w_t = torch.ones_like(score_tran) * mask[:, :, 0:1]
    w_t = w_t / (w_t.sum(dim=1, keepdim=True) + 1e-10)
    w_r = torch.ones_like(score_rot) * mask[:, :, 0:1]
    w_r = w_r / (w_r.sum(dim=1, keepdim=True) + 1e-10)
    ft_all = torch.cat((trans_feat0.unsqueeze(1), fused_tran.reshape(bs, n, 256)), dim=1)
    fr_all = torch.cat((rot_feat0.unsqueeze(1), fused_rot.reshape(bs, n, 256)), dim=1)
    if m > 1:
        ft_avg = torch.sum(ft_all * w_t, dim=1)
        fr_avg = torch.sum(fr_all * w_r, dim=1)
    else:
        ft_avg = (fused_tran.reshape(bs, n, 256) * w_t[:, 1:] / w_t[:, 1:].sum(dim=1, keepdim=True)).sum(dim=1)
        fr_avg = (fused_rot.reshape(bs, n, 256) * w_r[:, 1:] / w_r[:, 1:].sum(dim=1, keepdim=True)).sum(dim=1)
    rot_avg = F.normalize(linear(sd, "rots", fr_avg), dim=-1, p=2)
    tran_avg = linear(sd, "trans", ft_avg)
    out.update(pred_trans_avg=tran_avg, pred_rot_avg=rot_avg)
    if m <= 1:   # :1068-1075
        out.update(pred_trans=tran_avg, pred_rot=rot_avg)
        return out
    sel_rot = sel_tran = -1
    if out_cam_type == "avg-all":
        rot_fin, tran_fin = rot_avg, tran_avg
    elif out_cam_type == "soft":
        rot_fin = F.normalize(linear(sd, "rots", torch.sum(fr_all * score_rot, dim=1)), dim=-1, p=2)
        tran_fin = linear(sd, "trans", torch.sum(ft_all * score_tran, dim=1))
    elif out_cam_type == "min-cost":
        sel_rot = int(dL2_sum[0, :m + 1].argmin())
        sel_tran = int(dl2_sum[0, :m + 1].argmin())
        rot_fin, tran_fin = rots_all[0:1, sel_rot], trans_all[0:1, sel_tran]
    elif out_cam_type == "max-score":
        sel_rot = int(score_rot[0, :m + 1, 0].argmax())
        sel_tran = int(score_tran[0, :m + 1, 0].argmax())
        rot_fin, tran_fin = rots_all[0:1, sel_rot], trans_all[0:1, sel_tran]
    else:
        raise ValueError(out_cam_type)
    out.update(pred_trans=tran_fin, pred_rot=rot_fin, sel_rot=sel_rot, sel_tran=sel_tran,
               all_pred_trans=trans_all[0:1, :m + 1], all_pred_rots=rots_all[0:1, :m + 1],
               score_soft_rot=score_rot[0:1, :m + 1], score_soft_offset=score_tran[0:1, :m + 1],
               l2_dist=dl2_ori[0:1, :m + 1, :m], normal_dist=ang[0:1, :m + 1, :m],
               offset_dist=doff[0:1, :m + 1, :m], normal_l2_sum=dL2_sum[0:1, :m + 1],
               l2_sum=dl2_sum[0:1, :m + 1])
    return out


def ref_head(sd: SD, trans_feat0, rot_feat0, geo_global, geo_local, m, sig_seq, q0, t0, out_cam_type="soft"):
    """__inference_PlaneCamRefHead (camera_head.py:925-1115) for one pair."""
    fused_rot, fused_tran = hypothesis_features(sd, geo_global, sig_seq, rot_feat0, trans_feat0)
    return score_and_select(sd, fused_rot, fused_tran, rot_feat0, trans_feat0, geo_local, m, q0, t0, out_cam_type)


# ----------------------------------------------------------------------------------------------
# K10 — assignment pruning (camera_head.py:605-629)
# ----------------------------------------------------------------------------------------------
def prune_assignment(assign, params1, params2, q, t):
    p2w = warp_to_global(params2, pose_n=1)[:, 0]
    off2 = torch.norm(p2w, dim=2, keepdim=True, p=2)
    n2 = F.normalize(p2w, dim=-1, p=2)
    p1r = warp_to_global(params1, q.unsqueeze(1), t.unsqueeze(1) * 0.)[:, 0]
    n1r = F.normalize(p1r, dim=-1, p=2)
    nd = torch.acos(torch.clamp(torch.bmm(n1r, n2.transpose(1, 2)), -1, 1)) / math.pi * 180.
    p1rt = warp_to_global(params1, q.unsqueeze(1), t.unsqueeze(1))[:, 0]
    off1 = torch.norm(p1rt, dim=2, keepdim=True, p=2)
    n1rt = F.normalize(p1rt, dim=-1, p=2)
    nTn = torch.bmm(n1rt, n2.transpose(1, 2))
    od = torch.abs(off1 - off2.transpose(1, 2))
    od[nTn < 0] = torch.abs(off1 + off2.transpose(1, 2))[nTn < 0]
    od = torch.clamp(od, min=1e-4, max=10)
    return assign * ((nd < 45.) & (od < 1.)).to(assign.dtype)


# ----------------------------------------------------------------------------------------------
# whole head, one pair (camera_head.py:400-640), and the batch loop
# ----------------------------------------------------------------------------------------------
def inference_joint(sd: SD, msd: SD, feats1, feats2, params1, params2, app1, app2, *, num_queries,
                    out_cam_type="soft", match_threshold=0.2, hyp_pairs: Optional[torch.Tensor] = None,
                    initial_pose=None, sinkhorn_iters=200):
    """All tensors carry a leading batch dim of 1.  `hyp_pairs` [H,2] overrides the hypothesis list fed
    to the refinement head (the "P planes x H hypotheses" stress mapping of SURVEY.md §8(d)); the
    matcher still runs and its assignment is still what gets reported/pruned.  `initial_pose`=(t,q)
    skips K1 (stage set S3)."""
    out = {}
    if initial_pose is None:
        t_init, q_init, _, _ = pixel_camera_head(sd, feats1, feats2)
    else:
        t_init, q_init = initial_pose
    if q_init[0, 0] < 0:                      # :436-437
        q_init = -q_init
    out["camera_init"] = (t_init, q_init)
    q0, rot_feat0 = rot_rec_head(sd, q_init)  # :451
    t0, trans_feat0 = trans_rec_head(sd, t_init)
    out["camera_initRec"] = (t0, q0)
    cam = torch.cat([t0, q0], dim=-1)
    lsp = matching_head(msd, app1, app2, cam, params1, params2, sinkhorn_iters)   # :493-497
    out["log_scores_padded"] = lsp
    assign = get_assignment_matrix(lsp, match_threshold)                          # :501
    out["assignment_before"] = assign.clone()
    idx_pairs = torch.nonzero(assign[0]) if hyp_pairs is None else hyp_pairs
    local, glob, sig, m = geo_sequences(params1[0], params2[0], idx_pairs, num_queries, t0[0], q0[0])
    out.update(geo_local=local, geo_global=glob, sig_seq=sig, matched_num=m)
    r = ref_head(sd, trans_feat0, rot_feat0, glob[None], local[None], m, sig[None], q0, t0, out_cam_type)
    out["ref"] = r
    out["camera_avgRef0"] = (r["pred_trans_avg"], r["pred_rot_avg"])
    out["camera_softRef0"] = (r["pred_trans"], r["pred_rot"])
    out["camera"] = out["camera_softRef0"]
    q_ref, t_ref = r["pred_rot"], r["pred_trans"]
    if q_ref[0, 0] < 0:                       # :600-601
        q_ref = -q_ref
    out["assignment_after"] = prune_assignment(assign, params1, params2, q_ref, t_ref)
    return out


def inference_joint_batch(sd, msd, feats1, feats2, params1, params2, app1, app2, **kw):
    """Per-pair loop at bs=1 — the only batch mode the reference supports
    (meta_arch/siamese_planeTR.py:340; camera_modules.py:27)."""
    B = params1.shape[0]
    outs = []
    init = kw.pop("initial_pose", None)
    for b in range(B):
        f1 = None if feats1 is None else {k: v[b:b + 1] for k, v in feats1.items()}
        f2 = None if feats2 is None else {k: v[b:b + 1] for k, v in feats2.items()}
        ip = None if init is None else (init[0][b:b + 1], init[1][b:b + 1])
        outs.append(inference_joint(sd, msd, f1, f2, params1[b:b + 1], params2[b:b + 1],
                                    app1[b:b + 1], app2[b:b + 1], initial_pose=ip, **kw))
    return outs

"""SEM_SEG_HEAD of the reference — the DETR-style plane detector `PlaneTRHead` (first half) — on the
tensor-core engine: same registry surface (`SEM_SEG_HEADS_REGISTRY`, `build_planeTR_head(cfg, shape)`), constructor, state-dict
names and return value as modeling/planeTR_net/planeTR_head.py:23-192 (with the layers of modeling/transformer/transformer.py
and the sine position encoding of position_encoding.py), inference only.

    output, query_feat = head(features)          # features: {'res2'..'res5'} NCHW fp32, or a backbone.PlaneFeatures
    output = {'pred_logits' [b,NQ,2], 'pred_mask_logits' [b,NQ,H/4,W/4], 'pred_params' [b,NQ,3],
              'pixel_centers' [b,2,H/4,W/4], 'pred_centers' [b,NQ,2]};  query_feat = hs[-1] [b,NQ,256]

How it runs: every nn.Linear / 1x1 convolution is a GEMM on 16-bit hi/lo planes (`nsac_gemm_split*`, 3 passes ~
fp32; BatchNorm(eval) folded into the weights), the attention core is `nsac_attention_tiled` (exact fp32 softmax), and one
row kernel (`nsac_row_op`) does each residual add + LayerNorm + `with_pos_embed` in a single pass.  Two algebraic moves, both
exact in real arithmetic: (1) the top-down 1x1 convolutions + BatchNorm run BEFORE the bilinear 2x upsampling they follow in the
reference (both are per-pixel linear / affine, so they commute: 4x fewer GEMM rows), the upsampling + ReLU + lateral add is one
kernel; (2) the pixel-embedding convolution is folded into per-image plane weights, mask_logits = (plane_emb . W_pix) . p +
plane_emb . b_pix, so the [b,256,H/4,W/4] pixel-embedding map is never materialised.  Only the last decoder layer's output is
computed through the heads (the reference stacks all six for deep supervision and uses hs[-1] at inference).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn

from . import ops
from .compat import Registry, ShapeSpec

__all__ = ["PlaneTRHead", "SEM_SEG_HEADS_REGISTRY", "build_planeTR_head"]

SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")


def build_planeTR_head(cfg, shape):
    """planeTR_head.py:19-21."""
    return SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.NAME)(cfg, shape)


# ---------------------------------------------------------------------------------------------------
# parameter containers (names = reference state-dict keys; evaluated by the kernels, never called)
# ---------------------------------------------------------------------------------------------------
class _EncoderLayer(nn.Module):
    """transformer.py:142-168 (TransformerEncoderLayer)."""

    def __init__(self, d, nhead, ff):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=0.1)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)


class _DecoderLayer(nn.Module):
    """transformer.py:225-254 (TransformerDecoderLayer)."""

    def __init__(self, d, nhead, ff):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=0.1)
        self.multihead_attn = nn.MultiheadAttention(d, nhead, dropout=0.1)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)


class _Stack(nn.Module):
    def __init__(self, make_layer, num_layers, d):
        super().__init__()
        self.layers = nn.ModuleList([make_layer() for _ in range(num_layers)])
        self.norm = nn.LayerNorm(d)


class MLP(nn.Module):
    """planeTR_head.py:194-205."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


def conv_bn_relu(in_dim, out_dim, k=1, pad=0):
    """planeTR_head.py:206-212."""
    return nn.Sequential(nn.Conv2d(in_dim, out_dim, (k, k), padding=pad, bias=False), nn.BatchNorm2d(out_dim), nn.ReLU(inplace=True))


class top_down(nn.Module):
    """planeTR_head.py:215-238 (parameters only)."""

    def __init__(self, in_channels, channel, m_dim):
        super().__init__()
        self.up_conv3, self.up_conv2, self.up_conv1 = (conv_bn_relu(channel, channel, 1) for _ in range(3))
        self.c4_conv = conv_bn_relu(in_channels[3], channel, 1)
        self.c3_conv = conv_bn_relu(in_channels[2], channel, 1)
        self.c2_conv = conv_bn_relu(in_channels[1], channel, 1)
        self.c1_conv = conv_bn_relu(in_channels[0], channel, 1)
        self.m_conv_dict = nn.ModuleDict({"m4": conv_bn_relu(m_dim, channel)})


def sine_position_table(h: int, w: int, num_pos_feats: int, temperature: float = 10000.0) -> torch.Tensor:
    """PositionEmbeddingSine(num_pos_feats, normalize=True) for an un-masked h x w grid (position_encoding.py:29-52) as a table
    [h*w, 2*num_pos_feats] (row = y*w + x; first half = y terms, second half = x terms; sin on even, cos on odd channels)."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32) / (h + eps) * scale
    x = torch.arange(1, w + 1, dtype=torch.float32) / (w + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):
        p = v[:, None] / dim_t
        return torch.stack((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=2).flatten(1)
    py, px = enc(y), enc(x)
    return torch.cat((py[:, None, :].expand(h, w, -1), px[None, :, :].expand(h, w, -1)), dim=2).reshape(h * w, 2 * num_pos_feats).contiguous()


@SEM_SEG_HEADS_REGISTRY.register()
class PlaneTRHead(nn.Module):
    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        self.cfg = cfg
        S = cfg.MODEL.SEM_SEG_HEAD
        self.num_classes = S.NUM_CLASSES
        self.backbone_channels = [v.channels for v in input_shape.values()]
        self.param_on, self.center_on, self.depth_on = S.PARAM_ON, S.CENTER_ON, cfg.MODEL.DEPTH_ON
        if self.depth_on:
            raise NotImplementedError("MODEL.DEPTH_ON (top_down_depth / pixel_depth) is not used by the inference configs")
        self.hidden_dim, self.num_queries, self.nheads = S.HIDDEN_DIM, S.NUM_OBJECT_QUERIES, S.NHEADS
        self.enc_layers, self.dec_layers = S.ENC_LAYERS, S.DEC_LAYERS
        self.plane_embedding_dim = S.MASK_DIM
        self.channel = 256
        d = self.hidden_dim
        assert d % self.nheads == 0 and d // self.nheads == 32, "the attention kernel is built for head dim 32 (HIDDEN_DIM 256, NHEADS 8)"
        self.input_proj = nn.Conv2d(self.backbone_channels[-1], d, kernel_size=1)
        self.context_SA = _Stack(lambda: _EncoderLayer(d, self.nheads, 1024), self.enc_layers, d)
        self.query_embed = nn.Embedding(self.num_queries, d)
        self.context2plane_decoder = _Stack(lambda: _DecoderLayer(d, self.nheads, 1024), self.dec_layers, d)
        self.top_down = top_down(self.backbone_channels, self.channel, d)
        self.plane_embedding = MLP(d, d, self.plane_embedding_dim, 3)
        self.pixel_embedding = nn.Conv2d(self.channel, self.plane_embedding_dim, (1, 1), padding=0)
        self.plane_prob = nn.Linear(d, self.num_classes + 1)
        if self.param_on:
            self.plane_param = MLP(d, d, 3, 3)
        if self.center_on:
            self.plane_center = MLP(d, d, 2, 3)
            self.pixel_plane_center = nn.Conv2d(self.channel, 2, (1, 1), padding=0)
        self._packed = None
        self._pos = {}
        self.tc_passes = 3
        self.eval()

    # ------------------------------------------------------------------ weight packing
    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed, self._pos = None, {}
        return super()._apply(fn, *a, **k)

    def prepare(self):
        ver = ops.weights_version(self)
        if self._packed is not None and self._packed_version == ver:
            return self._packed
        d = self.hidden_dim
        sw = lambda w: ops.split_weight(w.detach().contiguous())
        c = lambda t: t.detach().contiguous()
        with torch.no_grad():
            pk = {"proj": (sw(self.input_proj.weight.reshape(d, -1)), c(self.input_proj.bias))}

            def attn(m, fused_qk):
                W, b = m.in_proj_weight.detach(), m.in_proj_bias.detach()
                out = {"v": (sw(W[2 * d:]), c(b[2 * d:])), "o": (sw(m.out_proj.weight), c(m.out_proj.bias))}
                if fused_qk:
                    out["qk"] = (sw(W[:2 * d]), c(b[:2 * d]))
                else:
                    out["q"], out["k"] = (sw(W[:d]), c(b[:d])), (sw(W[d:2 * d]), c(b[d:2 * d]))
                return out
            ffn = lambda l: {"l1": (sw(l.linear1.weight), c(l.linear1.bias)), "l2": (sw(l.linear2.weight), c(l.linear2.bias))}
            ln = lambda n: (c(n.weight), c(n.bias), float(n.eps))
            pk["enc"] = [{"sa": attn(l.self_attn, True), **ffn(l), "n1": ln(l.norm1), "n2": ln(l.norm2)} for l in self.context_SA.layers]
            pk["enc_norm"] = ln(self.context_SA.norm)
            pk["dec"] = [{"sa": attn(l.self_attn, True), "ca": attn(l.multihead_attn, False), **ffn(l), "n1": ln(l.norm1),
                          "n2": ln(l.norm2), "n3": ln(l.norm3)} for l in self.context2plane_decoder.layers]
            pk["dec_norm"] = ln(self.context2plane_decoder.norm)
            pk["query"] = c(self.query_embed.weight)

            def cbr(seq):     # conv (no bias) + BatchNorm2d(eval) folded: y = conv(x) * s + (beta - mean * s)
                conv, bn = seq[0], seq[1]
                s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
                w = conv.weight.detach().reshape(conv.weight.shape[0], -1) * s[:, None]
                return sw(w), (bn.bias.detach() - bn.running_mean * s).contiguous()
            td = self.top_down
            for name, seq in (("c4", td.c4_conv), ("c3", td.c3_conv), ("c2", td.c2_conv), ("c1", td.c1_conv), ("m4", td.m_conv_dict["m4"]),
                              ("up3", td.up_conv3), ("up2", td.up_conv2), ("up1", td.up_conv1)):
                pk["td." + name] = cbr(seq)
            mlp = lambda m: [(sw(l.weight), c(l.bias)) if l.weight.shape[0] >= 8 else (c(l.weight), c(l.bias)) for l in m.layers]
            pk["plane_embedding"] = mlp(self.plane_embedding)
            wpix = self.pixel_embedding.weight.detach().reshape(self.plane_embedding_dim, self.channel)       # [c_out, c_in]
            pk["pix_wT"] = sw(wpix.t())                                     # rows = c_in, K = c_out: W'[q, c_in] = sum_c pe[q, c] Wpix[c, c_in]
            pk["pix_b"] = c(self.pixel_embedding.bias).reshape(1, -1)
            pk["plane_prob"] = (c(self.plane_prob.weight), c(self.plane_prob.bias))
            if self.param_on:
                pk["plane_param"] = mlp(self.plane_param)
            if self.center_on:
                pk["plane_center"] = mlp(self.plane_center)
                w8 = torch.zeros(8, self.channel, device=wpix.device)
                w8[:2] = self.pixel_plane_center.weight.detach().reshape(2, self.channel)
                b8 = torch.zeros(8, device=wpix.device)
                b8[:2] = self.pixel_plane_center.bias.detach()
                pk["pix_center"] = (sw(w8), b8)
        self._packed, self._packed_version = pk, ver
        return pk

    def _pos_table(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._pos:
            self._pos[key] = sine_position_table(h, w, self.hidden_dim // 2).to(device)
        return self._pos[key]

    # ------------------------------------------------------------------ pieces
    def _run_mlp(self, layers, xp, x32):
        """3-layer MLP (ReLU between) on planes `xp` / fp32 `x32` -> fp32 [rows, out]."""
        P = self.tc_passes
        for i, (w, b) in enumerate(layers):
            last = i == len(layers) - 1
            act = ops.ACT_NONE if last else ops.ACT_RELU
            if isinstance(w, ops.Split):
                x32, xp = ops.gemm_tc(xp, w, b, act, P, want_f32=True, want_split=not last)
            else:                       # narrow output (2 / 3 columns): exact-fp32 CUDA-core linear
                x32 = ops.linear(x32, w, b, act)
        return x32

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, features):
        from .backbone import PlaneFeatures
        pk = self.prepare()
        P, d, NQ = self.tc_passes, self.hidden_dim, self.num_queries
        if isinstance(features, PlaneFeatures):
            lv = {k: features[k] for k in ("res2", "res3", "res4", "res5")}
            N = features.num_images
        else:
            N = features["res5"].shape[0]
            lv = {}
            for k in ("res2", "res3", "res4", "res5"):
                f = features[k]
                lv[k] = (ops.nchw_to_planes(f.float()), f.shape[2], f.shape[3])
        (c1p, h1, w1), (c2p, h2, w2), (c3p, h3, w3), (c4p, h4, w4) = lv["res2"], lv["res3"], lv["res4"], lv["res5"]
        dev = c4p.hi.device
        T, R, Q = h4 * w4, N * h4 * w4, N * NQ
        pos = self._pos_table(h4, w4, dev)
        mk = lambda rows, cols: torch.empty(rows, cols, device=dev, dtype=torch.float32)
        sp = lambda rows, cols=d: ops.Split.empty(rows, cols, dev)

        # ---- context projection + self-attention encoder (planeTR_head.py:124-131; post-norm layers, transformer.py:170-185)
        src, srcp = ops.gemm_tc(c4p, *pk["proj"], ops.ACT_NONE, P, want_f32=True, want_split=True)
        _, _, srcpos = ops.row_op(src, pos=pos, T=T, want_pos_split=True)                       # q = k = src + pos
        qkv = mk(R, 3 * d)
        for lw in pk["enc"]:
            ops.gemm_tc(srcpos, *lw["sa"]["qk"], ops.ACT_NONE, P, out_f32=qkv[:, :2 * d])
            ops.gemm_tc(srcp, *lw["sa"]["v"], ops.ACT_NONE, P, out_f32=qkv[:, 2 * d:])          # value = src (no position)
            att = ops.attention_tiled(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], N, T, T, H=self.nheads, out_split=sp(R))
            a32, _ = ops.gemm_tc(att, *lw["sa"]["o"], ops.ACT_NONE, P)
            src, srcp, _ = ops.row_op(src, y=a32, ln=lw["n1"], want_f32=True, want_split=True)  # norm1(src + attn)
            _, hp = ops.gemm_tc(srcp, *lw["l1"], ops.ACT_RELU, P, want_f32=False, want_split=True)
            f32, _ = ops.gemm_tc(hp, *lw["l2"], ops.ACT_NONE, P)
            src, srcp, srcpos = ops.row_op(src, y=f32, ln=lw["n2"], pos=pos, T=T, want_f32=True, want_split=True, want_pos_split=True)
        mem, memp, mempos = ops.row_op(src, ln=pk["enc_norm"], pos=pos, T=T, want_f32=True, want_split=True, want_pos_split=True)

        # ---- plane decoder (:134-139; pre-norm layers, transformer.py:284-311), tgt = 0, query_pos = query_embed
        qe = pk["query"]
        tgt = torch.zeros(Q, d, device=dev)
        _, t2p, t2qp = ops.row_op(tgt, ln=pk["dec"][0]["n1"], pos=qe, T=NQ, want_split=True, want_pos_split=True)
        qkv_d, kv_m = mk(Q, 3 * d), mk(R, 2 * d)
        hs32 = hsp = None
        for i, lw in enumerate(pk["dec"]):
            ops.gemm_tc(t2qp, *lw["sa"]["qk"], ops.ACT_NONE, P, out_f32=qkv_d[:, :2 * d])
            ops.gemm_tc(t2p, *lw["sa"]["v"], ops.ACT_NONE, P, out_f32=qkv_d[:, 2 * d:])
            att = ops.attention_tiled(qkv_d[:, :d], qkv_d[:, d:2 * d], qkv_d[:, 2 * d:], N, NQ, NQ, H=self.nheads, out_split=sp(Q))
            a32, _ = ops.gemm_tc(att, *lw["sa"]["o"], ops.ACT_NONE, P)
            _, t2p, t2qp = ops.row_op(tgt, y=a32, sum_out=tgt, ln=lw["n2"], pos=qe, T=NQ, want_split=True, want_pos_split=True)
            q32, _ = ops.gemm_tc(t2qp, *lw["ca"]["q"], ops.ACT_NONE, P)                          # query = norm2(tgt) + query_pos
            ops.gemm_tc(mempos, *lw["ca"]["k"], ops.ACT_NONE, P, out_f32=kv_m[:, :d])            # key = memory + pos
            ops.gemm_tc(memp, *lw["ca"]["v"], ops.ACT_NONE, P, out_f32=kv_m[:, d:])              # value = memory
            att = ops.attention_tiled(q32, kv_m[:, :d], kv_m[:, d:], N, NQ, T, H=self.nheads, out_split=sp(Q))
            a32, _ = ops.gemm_tc(att, *lw["ca"]["o"], ops.ACT_NONE, P)
            _, t2p, _ = ops.row_op(tgt, y=a32, sum_out=tgt, ln=lw["n3"], want_split=True)
            _, hp = ops.gemm_tc(t2p, *lw["l1"], ops.ACT_RELU, P, want_f32=False, want_split=True)
            f32, _ = ops.gemm_tc(hp, *lw["l2"], ops.ACT_NONE, P)
            if i + 1 < len(pk["dec"]):
                _, t2p, t2qp = ops.row_op(tgt, y=f32, sum_out=tgt, ln=pk["dec"][i + 1]["n1"], pos=qe, T=NQ, want_split=True, want_pos_split=True)
            else:
                hs32, hsp, _ = ops.row_op(tgt, y=f32, ln=pk["dec_norm"], want_f32=True, want_split=True)   # hs[-1] = norm(output)

        # ---- pixel decoder (top_down, :240-252): conv + BN at the LOW resolution, then upsample + ReLU + lateral add
        a4, _ = ops.gemm_tc(c4p, *pk["td.c4"], ops.ACT_RELU, P)
        b4, _ = ops.gemm_tc(memp, *pk["td.m4"], ops.ACT_RELU, P)
        _, p4p, _ = ops.row_op(a4, y=b4, want_split=True)
        u3, _ = ops.gemm_tc(p4p, *pk["td.up3"], ops.ACT_NONE, P)
        l3, _ = ops.gemm_tc(c3p, *pk["td.c3"], ops.ACT_RELU, P)
        _, p3p = ops.upsample2x_relu_add(u3, l3, N, h4, w4)
        u2, _ = ops.gemm_tc(p3p, *pk["td.up2"], ops.ACT_NONE, P)
        l2, _ = ops.gemm_tc(c2p, *pk["td.c2"], ops.ACT_RELU, P)
        _, p2p = ops.upsample2x_relu_add(u2, l2, N, h3, w3)
        u1, _ = ops.gemm_tc(p2p, *pk["td.up1"], ops.ACT_NONE, P)
        l1, _ = ops.gemm_tc(c1p, *pk["td.c1"], ops.ACT_RELU, P)
        _, p1p = ops.upsample2x_relu_add(u1, l1, N, h2, w2)                                     # p_context [N*h1*w1, 256] as planes
        assert (2 * h2, 2 * w2) == (h1, w1) and (2 * h3, 2 * w3) == (h2, w2) and (2 * h4, 2 * w4) == (h3, w3), \
            "PlaneTRHead needs feature maps whose sizes double level to level (input size divisible by 32)"

        # ---- heads (:148-160).  mask logits: (plane_emb . W_pix) . p + plane_emb . b_pix, one GEMM per image with M = NQ rows
        pe = self._run_mlp(pk["plane_embedding"], hsp, hs32)
        wq32, wqp = ops.gemm_tc(ops.split(pe), pk["pix_wT"], None, ops.ACT_NONE, P, want_f32=True, want_split=True)
        rowb = ops.linear(pe, pk["pix_b"]).reshape(Q)
        HW = h1 * w1
        mask = torch.empty(N, NQ, h1, w1, device=dev)
        for n in range(N):
            ops.gemm_tc_rowbias(wqp.rows_view(n * NQ, (n + 1) * NQ), p1p.rows_view(n * HW, (n + 1) * HW), rowb[n * NQ:(n + 1) * NQ],
                                mask[n].view(NQ, HW), P)
        out = {"pred_logits": ops.linear(hs32, *pk["plane_prob"]).view(N, NQ, -1), "pred_mask_logits": mask}
        if self.param_on:
            out["pred_params"] = self._run_mlp(pk["plane_param"], hsp, hs32).view(N, NQ, 3)
        if self.center_on:
            out["pred_centers"] = torch.sigmoid(self._run_mlp(pk["plane_center"], hsp, hs32)).view(N, NQ, 2)
            pc, _ = ops.gemm_tc(p1p, *pk["pix_center"], ops.ACT_NONE, P)
            out["pixel_centers"] = torch.sigmoid(pc[:, :2]).view(N, h1, w1, 2).permute(0, 3, 1, 2).contiguous()
        return out, hs32.view(N, NQ, d)

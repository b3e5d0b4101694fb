"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch, functional, from a state dict) of the reference's plane detector
`PlaneTRHead` (first half): /root/reference/NopeSAC_Net/modeling/planeTR_net/planeTR_head.py:116-192 with
the DETR-style layers of modeling/transformer/transformer.py and the sine position encoding of position_encoding.py.

Pinned: bit-level against the LIVE reference module imported from /root/reference (oracle/ref_planetr_loader.py,
tests/test_oracle_planetr.py) and against tests/golden/planetr_*.pt generated from it (tests/golden/make_planetr_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; nothing under nopesac_b200/ does.
Inference only (eval mode: dropout is the identity, BatchNorm uses its running statistics)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def position_embedding_sine(b: int, h: int, w: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """position_encoding.py:29-52 with normalize=True, scale=2*pi, no mask -> [b, 2*num_pos_feats, h, w]."""
    not_mask = torch.ones(b, h, w, dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def _mha(sd: SD, p: str, q, k, v, nhead: int):
    """nn.MultiheadAttention(d_model, nhead) forward on [L, N, E] tensors (transformer.py:153, 238-239)."""
    E = q.shape[-1]
    return F.multi_head_attention_forward(q, k, v, E, nhead, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"], None, None, False,
                                          0.0, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], training=False,
                                          need_weights=False)[0]


def _ln(sd: SD, p: str, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def encoder(sd: SD, p: str, src, pos, num_layers: int, nhead: int):
    """TransformerEncoder of post-norm layers + final norm (transformer.py:76-100, forward_post :170-185; planeTR_head.py:78-82)."""
    out = src
    for i in range(num_layers):
        lp = f"{p}.layers.{i}"
        q = k = out + pos
        out = _ln(sd, lp + ".norm1", out + _mha(sd, lp + ".self_attn", q, k, out, nhead))
        out = _ln(sd, lp + ".norm2", out + _lin(sd, lp + ".linear2", F.relu(_lin(sd, lp + ".linear1", out))))
    return _ln(sd, p + ".norm", out)


def decoder_last(sd: SD, p: str, tgt, memory, pos, query_pos, num_layers: int, nhead: int):
    """TransformerDecoder of pre-norm layers (forward_pre :284-311); returns norm(output of the last layer) = hs[-1]
    (transformer.py:103-150 with return_intermediate: the last stacked element is the final norm)."""
    out = tgt
    for i in range(num_layers):
        lp = f"{p}.layers.{i}"
        t2 = _ln(sd, lp + ".norm1", out)
        q = k = t2 + query_pos
        out = out + _mha(sd, lp + ".self_attn", q, k, t2, nhead)
        t2 = _ln(sd, lp + ".norm2", out)
        out = out + _mha(sd, lp + ".multihead_attn", t2 + query_pos, memory + pos, memory, nhead)
        t2 = _ln(sd, lp + ".norm3", out)
        out = out + _lin(sd, lp + ".linear2", F.relu(_lin(sd, lp + ".linear1", t2)))
    return _ln(sd, p + ".norm", out)


def _mlp(sd: SD, p: str, x, n: int = 3):
    for i in range(n):
        x = _lin(sd, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def _conv_bn_relu(sd: SD, p: str, x):
    """conv_bn_relu (planeTR_head.py:206-212): 1x1 conv without bias + BatchNorm2d (eval) + ReLU."""
    x = F.conv2d(x, sd[p + ".0.weight"])
    x = F.batch_norm(x, sd[p + ".1.running_mean"], sd[p + ".1.running_var"], sd[p + ".1.weight"], sd[p + ".1.bias"], False, 0.1, 1e-5)
    return F.relu(x)


def top_down(sd: SD, p: str, feats, memory):
    """top_down.forward (planeTR_head.py:240-252)."""
    c1, c2, c3, c4 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    p4 = _conv_bn_relu(sd, p + ".c4_conv", c4) + _conv_bn_relu(sd, p + ".m_conv_dict.m4", memory)
    p3 = _conv_bn_relu(sd, p + ".up_conv3", up(p4)) + _conv_bn_relu(sd, p + ".c3_conv", c3)
    p2 = _conv_bn_relu(sd, p + ".up_conv2", up(p3)) + _conv_bn_relu(sd, p + ".c2_conv", c2)
    return _conv_bn_relu(sd, p + ".up_conv1", up(p2)) + _conv_bn_relu(sd, p + ".c1_conv", c1)


def plane_tr_head(sd: SD, features: Dict[str, torch.Tensor], nhead: int = 8, enc_layers: int = 6, dec_layers: int = 6,
                  param_on: bool = True, center_on: bool = True) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """PlaneTRHead.forward (planeTR_head.py:116-192), eval mode, DEPTH_ON False -> (output dict, hs[-1] [b, NQ, hidden])."""
    c1, c2, c3, c4 = features["res2"], features["res3"], features["res4"], features["res5"]
    hidden = sd["input_proj.weight"].shape[0]
    pos_map = position_embedding_sine(c4.shape[0], c4.shape[2], c4.shape[3], hidden // 2).to(c4.dtype)        # :124
    feat_map = F.conv2d(c4, sd["input_proj.weight"], sd["input_proj.bias"])                                    # :125
    bs, _, hc, wc = feat_map.shape
    feat_seq = feat_map.flatten(2).permute(2, 0, 1)
    pos_seq = pos_map.flatten(2).permute(2, 0, 1)
    memory = encoder(sd, "context_SA", feat_seq, pos_seq, enc_layers, nhead)                                   # :131
    query_embed = sd["query_embed.weight"].unsqueeze(1).repeat(1, bs, 1)                                       # :134
    hs_last = decoder_last(sd, "context2plane_decoder", torch.zeros_like(query_embed), memory, pos_seq, query_embed, dec_layers,
                           nhead).transpose(0, 1)                                                              # :136-139 -> [b, NQ, hidden]
    mem_map = memory.permute(1, 2, 0).reshape(bs, hidden, hc, wc)                                              # :143
    p_context = top_down(sd, "top_down", (c1, c2, c3, c4), mem_map)                                            # :144
    plane_embedding = _mlp(sd, "plane_embedding", hs_last)                                                     # :148
    pixel_embedding = F.conv2d(p_context, sd["pixel_embedding.weight"], sd["pixel_embedding.bias"])            # :149
    out = {"pred_logits": _lin(sd, "plane_prob", hs_last),                                                     # :153
           "pred_mask_logits": torch.einsum("bqc,bchw->bqhw", plane_embedding, pixel_embedding)}               # :150
    if param_on:
        out["pred_params"] = _mlp(sd, "plane_param", hs_last)                                                  # :157
    if center_on:
        out["pred_centers"] = torch.sigmoid(_mlp(sd, "plane_center", hs_last))                                 # :159-160
        out["pixel_centers"] = torch.sigmoid(F.conv2d(p_context, sd["pixel_plane_center.weight"], sd["pixel_plane_center.bias"]))
    return out, hs_last

"""MATCHING_HEAD of the path — same registry surface, constructor, call signature, return value and
state-dict names as the reference `MatchingHead` (matching_net/matching_head.py:15-133,
transformer/gnn.py:46-138), computed by the CUDA kernels of libnopesac_b200.

    MATCHING_HEAD_REGISTRY.get("MatchingHead")(cfg)
    losses, log_scores_padded = head(planeApp1, planeApp2, matcher_inputCam, params1, params2)

Inference only (the training losses of the reference are out of scope).  In addition to the
reference return value the fused matcher kernel also yields the mutual-NN assignment matrix
(camera_modules.py:15-34); the camera head picks it up through `match()`.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import ops
from .compat import Registry

__all__ = ["build_matching_head", "MATCHING_HEAD_REGISTRY", "MatchingHead"]

MATCHING_HEAD_REGISTRY = Registry("MATCHING_HEAD")
MATCHING_HEAD_REGISTRY.__doc__ = "Registry for plane matching head"


def build_matching_head(cfg):
    # the reference ignores cfg.MODEL.MATCHING_HEAD.NAME (matching_head.py:20-21)
    return MATCHING_HEAD_REGISTRY.get("MatchingHead")(cfg)


class _EncoderLayerParams(nn.Module):
    """Parameter container with the names of gnn.py:56-71."""

    def __init__(self, d_model: int):
        super().__init__()
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(
            nn.Linear(d_model * 2, d_model * 2, bias=False), nn.ReLU(True), nn.Linear(d_model * 2, d_model, bias=False))
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class _GNNParams(nn.Module):
    def __init__(self, d_model: int, nhead: int, layer_names):
        super().__init__()
        self.d_model, self.nhead, self.layer_names = d_model, nhead, list(layer_names)
        self.layers = nn.ModuleList([_EncoderLayerParams(d_model) for _ in self.layer_names])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


@MATCHING_HEAD_REGISTRY.register()
class MatchingHead(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.offset_multiplier = cfg.MODEL.MATCHING_HEAD.OFFSET_MULTIPLIER
        self.normal_multiplier = cfg.MODEL.MATCHING_HEAD.NORMAL_MULTIPLIER
        self.gnn = _GNNParams(256, 8, ["self", "cross"] * 9)
        self.planeDesc_proj = nn.Conv1d(256, 256, kernel_size=1, bias=True)
        self.bin_score = torch.nn.Parameter(torch.tensor(1.0), requires_grad=True)
        self.sinkhorn_iterations = 200
        self.max_length = cfg.MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES
        self.mask_on = True
        self.planeApp_proj = nn.Conv1d(256, 256, kernel_size=1, bias=True)
        self.match_threshold = cfg.TEST.MATCHING_SCORE_THRESHOLD
        self._packed = None
        self.use_stage_entry = not os.environ.get("NSAC_PY_STAGES")    # nsac_match_forward; NSAC_PY_STAGES=1: same launches from Python
        self.tc_passes = 3     # MMA passes of the tensor-core layers (fp16 hi/lo planes, ~fp32)

    # ------------------------------------------------------------------ weight packing
    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def prepare(self):
        """Fused QKV / KV weight matrices (built once; invalidated by load_state_dict / .to())."""
        ver = ops.weights_version(self)
        if self._packed is None or self._packed_version != ver:
            self._packed_version = ver
            with torch.no_grad():
                pk = []
                for lyr in self.gnn.layers:
                    wq, wk, wv = lyr.q_proj.weight, lyr.k_proj.weight, lyr.v_proj.weight
                    pk.append({
                        "qkv": torch.cat([wq, wk, wv], 0).contiguous(),
                        "q": wq.detach().contiguous(),
                        "kv": torch.cat([wk, wv], 0).contiguous(),
                        "merge": lyr.merge.weight.detach().contiguous(),
                        "mlp0": lyr.mlp[0].weight.detach().contiguous(),
                        "mlp2": lyr.mlp[2].weight.detach().contiguous(),
                        "n1w": lyr.norm1.weight.detach(), "n1b": lyr.norm1.bias.detach(),
                        "n2w": lyr.norm2.weight.detach(), "n2b": lyr.norm2.bias.detach(),
                    })
                self._packed = {
                    "layers": pk,
                    "app_w": self.planeApp_proj.weight.detach().reshape(256, 256).contiguous(),
                    "app_b": self.planeApp_proj.bias.detach().contiguous(),
                    "desc_w": self.planeDesc_proj.weight.detach().reshape(256, 256).contiguous(),
                    "desc_b": self.planeDesc_proj.bias.detach().contiguous(),
                }
        return self._packed

    # ------------------------------------------------------------------ GNN (gnn.py:73-138)
    def prepare_tc(self):
        """fp16 hi/lo planes of every GNN / projection weight for the tensor-core engine."""
        pk = self.prepare()
        if "tc" not in pk:
            with torch.no_grad():
                pk["tc"] = [{k: ops.split_weight(w[k]) for k in ("qkv", "q", "kv", "merge", "mlp0", "mlp2")} for w in pk["layers"]]
                pk["app_ws"], pk["desc_ws"] = ops.split_weight(pk["app_w"]), ops.split_weight(pk["desc_w"])
        return pk

    def match_weights(self):
        """`nsac_match_weights` for `ops.match_forward` (borrowed pointers into this weight version's packed planes)."""
        pk = self.prepare_tc()
        if "match_struct" not in pk:
            from . import _lib
            W = _lib.MatchWeights()
            layers = (_lib.GnnLayer * len(pk["layers"]))()
            for g, w, ws, name in zip(layers, pk["layers"], pk["tc"], self.gnn.layer_names):
                for k in ("qkv", "q", "kv", "merge", "mlp0", "mlp2"):
                    setattr(g, k, ops.tc_layer(ws[k], None))
                g.n1w, g.n1b, g.n2w, g.n2b = (w[k].data_ptr() for k in ("n1w", "n1b", "n2w", "n2b"))
                g.self_attn = 1 if name == "self" else 0
            W.app_proj, W.desc_proj = ops.tc_layer(pk["app_ws"], pk["app_b"]), ops.tc_layer(pk["desc_ws"], pk["desc_b"])
            W.layers, W.num_layers = layers, len(pk["layers"])
            pk["bin_score"] = self.bin_score.detach().reshape(1).float().contiguous()
            W.bin_score = pk["bin_score"].data_ptr()
            W.offset_multiplier, W.normal_multiplier = float(self.offset_multiplier), float(self.normal_multiplier)
            W.sinkhorn_iterations, W.fmt, W.passes = int(self.sinkhorn_iterations), ops.SPLIT_F16, self.tc_passes
            pk["match_struct"], pk["match_struct.keep"] = W, layers
        return pk["match_struct"]

    @staticmethod
    def _layer(w, ws, X, Xp, xs, ss, B, L, S, self_attn: bool, P: int, kv_count=None):
        """X [rows,512] fp32 and Xp (its fp16 hi/lo planes): columns 0:256 hold the token features, 256:512 the
        message slot, so cat[x, message] (gnn.py:93) is free.  xs / ss = row ranges of the query / source stream.
        All six linears run on the tcgen05 engine; attention and LayerNorm emit the operand planes directly."""
        x, xp = X[xs[0]:xs[1]], Xp.rows_view(xs[0], xs[1])
        if self_attn:
            qkv, _ = ops.gemm_tc(xp.cols(0, 256), ws["qkv"], passes=P)
            q, k, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
        else:
            q, _ = ops.gemm_tc(xp.cols(0, 256), ws["q"], passes=P)
            kv, _ = ops.gemm_tc(Xp.rows_view(ss[0], ss[1]).cols(0, 256), ws["kv"], passes=P)
            k, v = kv[:, :256], kv[:, 256:]
        msgp = ops.Split.empty(x.shape[0], 256, x.device)
        ops.attention(q, k, v, B, L, S, out_split=msgp, want_f32=False, kv_count=kv_count)
        msg, _ = ops.gemm_tc(msgp, ws["merge"], passes=P)
        ops.layernorm(msg, w["n1w"], w["n1b"], out=x[:, 256:], out_split=xp.cols(256, 512))     # message slot
        _, hp = ops.gemm_tc(xp, ws["mlp0"], act=ops.ACT_RELU, passes=P, want_f32=False, want_split=True)
        msg, _ = ops.gemm_tc(hp, ws["mlp2"], passes=P)
        ops.layernorm(msg, w["n2w"], w["n2b"], res=x[:, :256], out=x[:, :256], out_split=xp.cols(0, 256))  # x + norm2(.)

    def _descriptors(self, planeApp1, planeApp2, count1=None, count2=None):
        pk = self.prepare_tc()
        P = self.tc_passes
        B, n1, _ = planeApp1.shape
        n2 = planeApp2.shape[1]
        R0, R1 = B * n1, B * n2
        dev = planeApp1.device
        X = torch.empty(R0 + R1, 512, device=dev, dtype=torch.float32)
        Xp = ops.Split.empty(R0 + R1, 512, dev)
        app = ops.split(torch.cat([planeApp1.reshape(R0, 256), planeApp2.reshape(R1, 256)], 0))
        ops.gemm_tc(app, pk["app_ws"], pk["app_b"], passes=P, out_f32=X[:, :256], want_split=True, out_split=Xp.cols(0, 256))
        s0, s1 = (0, R0), (R0, R0 + R1)
        # 'self' layers apply the SAME weights to both views independently (gnn.py:128-130): with equal plane counts the two calls
        # are one call over 2B batch elements (rows of view 1 follow those of view 0) - half the launches of these layers
        both = (0, R0 + R1)
        count12 = None if count1 is None else torch.cat([count1, count2])
        for w, ws, name in zip(pk["layers"], pk["tc"], self.gnn.layer_names):
            if name == "self" and n1 == n2:
                self._layer(w, ws, X, Xp, both, both, 2 * B, n1, n1, True, P, count12)
            elif name == "self":
                self._layer(w, ws, X, Xp, s0, s0, B, n1, n1, True, P, count1)
                self._layer(w, ws, X, Xp, s1, s1, B, n2, n2, True, P, count2)
            else:
                self._layer(w, ws, X, Xp, s0, s1, B, n1, n2, False, P, count2)
                self._layer(w, ws, X, Xp, s1, s0, B, n2, n1, False, P, count1)     # sees the UPDATED feat0 (gnn.py:133-134)
        desc, _ = ops.gemm_tc(Xp.cols(0, 256), pk["desc_ws"], pk["desc_b"], passes=P)
        return desc[:R0].view(B, n1, 256), desc[R0:].view(B, n2, 256)

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def match(self, planeApp1, planeApp2, matcher_inputCam, parameters1_local, parameters2_local,
              match_threshold=None, normal_decay=1.0, offset_deacy=1.0, plane_count1=None, plane_count2=None):
        """-> (log_scores_padded [B,n1+1,n2+1], assignment [B,n1,n2]).
        Ragged batches: `plane_count1` / `plane_count2` (int32 [B] on the device, e.g. `PlaneLists.count`) say how many of the
        padded n1 / n2 rows of pair b are planes; padded tokens are never attended to (the all-valid case of the reference's
        `kv_mask`, gnn.py:31-34) and every pair solves its own transport problem — results in the top-left block, the rest
        -inf / 0.  No host synchronisation."""
        if matcher_inputCam is None:
            raise NotImplementedError("matcher_inputCam=None is a training-only branch of the reference")
        if normal_decay != 1.0 or offset_deacy != 1.0:
            raise NotImplementedError("decay factors other than 1.0 are never used by the reference (camera_head.py:490-497)")
        if (plane_count1 is None) != (plane_count2 is None):
            raise ValueError("plane_count1 and plane_count2 go together")
        thr = self.match_threshold if match_threshold is None else match_threshold
        if self.use_stage_entry:           # the whole forward behind ONE C call (nsac_match_forward, csrc/forward.cu)
            W = self.match_weights()
            W.passes, W.sinkhorn_iterations = self.tc_passes, int(self.sinkhorn_iterations)
            return ops.match_forward(W, planeApp1.float(), planeApp2.float(), parameters1_local, parameters2_local, matcher_inputCam,
                                     float(thr), plane_count1, plane_count2)
        d1, d2 = self._descriptors(planeApp1.float(), planeApp2.float(), plane_count1, plane_count2)
        return ops.match_sinkhorn_assign(d1, d2, parameters1_local, parameters2_local, matcher_inputCam,
                                         self.bin_score.detach(), float(self.offset_multiplier),
                                         float(self.normal_multiplier), self.sinkhorn_iterations, float(thr),
                                         count1=plane_count1, count2=plane_count2)

    def forward(self, planeApp1, planeApp2, matcher_inputCam, parameters1_local, parameters2_local,
                indices1=None, indices2=None, gt_corr_matrix=None, suffix="", normal_decay=1.0, offset_deacy=1.0):
        if self.training:
            raise NotImplementedError("nopesac_b200.MatchingHead is inference-only")
        lsp, _ = self.match(planeApp1, planeApp2, matcher_inputCam, parameters1_local, parameters2_local,
                            normal_decay=normal_decay, offset_deacy=offset_deacy)
        return {}, lsp

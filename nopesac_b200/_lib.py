"""ctypes binding of libnopesac_b200.so (include/nopesac_b200.h).  No fallback: if the library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

c_float_p = C.c_void_p   # device pointers travel as integers
c_int_p = C.c_void_p


class ScoreMLP(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")]


class TcLayer(C.Structure):
    """nsac_tc_layer (include/nopesac_b200.h)."""
    _fields_ = [("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("bias", C.c_void_p), ("N", C.c_int), ("K", C.c_int), ("ldw", C.c_int),
                ("w_scale", C.c_float)]


class RefineWeights(C.Structure):
    """nsac_refine_weights (include/nopesac_b200.h)."""
    _fields_ = [("geo0_w", C.c_void_p), ("geo0_b", C.c_void_p),
                ("geo_encoder", TcLayer * 5), ("geo_proj_s1", TcLayer * 3), ("decoder_rot", TcLayer * 6), ("geo_proj_s2", TcLayer * 3),
                ("decoder_tran", TcLayer * 6), ("decoder_rot2", TcLayer * 3), ("decoder_tran2", TcLayer * 3),
                ("rot2_w_init", C.c_void_p), ("rot2_b0", C.c_void_p), ("tran2_w_init", C.c_void_p), ("tran2_b0", C.c_void_p),
                ("rots_w", C.c_void_p), ("rots_b", C.c_void_p), ("trans_w", C.c_void_p), ("trans_b", C.c_void_p),
                ("score_pack", C.c_void_p), ("score_vecs_host", C.c_void_p),
                ("rot_mlp", C.POINTER(ScoreMLP)), ("tran_mlp", C.POINTER(ScoreMLP)), ("fmt", C.c_int), ("passes", C.c_int)]


class GnnLayer(C.Structure):
    """nsac_gnn_layer."""
    _fields_ = [("qkv", TcLayer), ("q", TcLayer), ("kv", TcLayer), ("merge", TcLayer), ("mlp0", TcLayer), ("mlp2", TcLayer),
                ("n1w", C.c_void_p), ("n1b", C.c_void_p), ("n2w", C.c_void_p), ("n2b", C.c_void_p), ("self_attn", C.c_int)]


class MatchWeights(C.Structure):
    """nsac_match_weights."""
    _fields_ = [("app_proj", TcLayer), ("desc_proj", TcLayer), ("layers", C.POINTER(GnnLayer)), ("num_layers", C.c_int),
                ("bin_score", C.c_void_p), ("offset_multiplier", C.c_float), ("normal_multiplier", C.c_float),
                ("sinkhorn_iterations", C.c_int), ("fmt", C.c_int), ("passes", C.c_int)]


class PixelWeights(C.Structure):
    """nsac_pixel_weights."""
    _fields_ = [("pd_layer_3", TcLayer), ("pd_layer_2", TcLayer), ("pd_layer_1", TcLayer), ("pd_mask_features", TcLayer),
                ("pd_adapter_2", TcLayer), ("pd_adapter_1", TcLayer), ("gn_w", C.c_void_p * 5), ("gn_b", C.c_void_p * 5),
                ("gn_groups", C.c_int), ("gn_eps", C.c_float), ("cb", TcLayer * 6), ("ct0", TcLayer),
                ("convs_trans", TcLayer * 5), ("convs_rots", TcLayer * 5),
                ("fc_trans_w", C.c_void_p), ("fc_trans_b", C.c_void_p), ("fc_rots_w", C.c_void_p), ("fc_rots_b", C.c_void_p),
                ("rot_emb0_w", C.c_void_p), ("rot_emb0_b", C.c_void_p), ("trans_emb0_w", C.c_void_p), ("trans_emb0_b", C.c_void_p),
                ("rot_emb", TcLayer * 5), ("trans_emb", TcLayer * 5),
                ("rots_w", C.c_void_p), ("rots_b", C.c_void_p), ("trans_w", C.c_void_p), ("trans_b", C.c_void_p),
                ("fmt", C.c_int), ("passes", C.c_int)]


class HeadWeights(C.Structure):
    """nsac_head_weights."""
    _fields_ = [("pixel", C.POINTER(PixelWeights)), ("match", C.POINTER(MatchWeights)), ("refine", C.POINTER(RefineWeights))]


class Bottleneck(C.Structure):
    """nsac_bottleneck."""
    _fields_ = [("conv1", TcLayer), ("conv2", TcLayer), ("conv3", TcLayer), ("shortcut", TcLayer), ("has_shortcut", C.c_int), ("stride", C.c_int)]


class BackboneWeights(C.Structure):
    """nsac_backbone_weights."""
    _fields_ = [("stem", TcLayer), ("blocks", C.POINTER(Bottleneck)), ("num_blocks", C.c_int), ("stage_blocks", C.c_int * 4),
                ("fmt", C.c_int), ("passes", C.c_int)]


_SIGNATURES = {
    "nsac_version": (C.c_int, []),
    "nsac_last_error": (C.c_char_p, []),
    "nsac_linear": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int,
                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nsac_gemm_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, c_float_p, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_gemm_split_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, c_float_p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_int, c_float_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_gemm_split_rowbias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, c_float_p, c_float_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_row_op": (C.c_int, [c_float_p, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_float, C.c_int, c_float_p, C.c_int,
                              c_float_p, C.c_int, c_float_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int, C.c_int, C.c_void_p]),
    "nsac_attention_tiled": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nsac_upsample2x_relu_add": (C.c_int, [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "nsac_plane_overflow": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "nsac_split16": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_void_p]),
    "nsac_conv3x3_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_float_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_conv3x3_split_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_float_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_conv1x1_split_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_float_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nsac_nchw_to_planes": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsac_groupnorm_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "nsac_groupnorm_nhwc": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, C.c_float,
                                      C.c_int, c_float_p, C.c_int, c_float_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "nsac_maxpool2_planes": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "nsac_corr_softmax": (C.c_int, [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "nsac_im2col3x3_planes": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "nsac_layernorm": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, c_float_p,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nsac_attention": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p]),
    "nsac_attention_ragged": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        c_int_p, C.c_void_p]),
    "nsac_match_sinkhorn_assign_ragged": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                                    C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int,
                                                    C.c_int, C.c_int, c_int_p, c_int_p, c_float_p, c_float_p, C.c_void_p]),
    "nsac_match_sinkhorn_assign": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                             C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int,
                                             C.c_int, C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "nsac_geo_sequence": (C.c_int, [c_float_p, c_float_p, c_float_p, c_int_p, C.c_int, c_float_p, c_float_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p,
                                    c_float_p, c_int_p, c_int_p, C.c_void_p]),
    "nsac_pose_heads": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int,
                                  C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "nsac_score_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "nsac_score_aggregate": (C.c_int, [c_float_p] * 9 + [c_int_p, C.POINTER(ScoreMLP), C.POINTER(ScoreMLP)] +
                             [c_float_p] * 4 + [C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p,
                                                c_int_p, c_float_p, C.c_void_p, C.c_void_p]),
    "nsac_score_pack_bytes": (C.c_size_t, [C.c_int]),
    "nsac_score_pack": (C.c_int, [C.POINTER(ScoreMLP), C.POINTER(ScoreMLP), C.c_int, C.c_void_p, C.c_void_p]),
    "nsac_score_tc_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "nsac_score_aggregate_tc": (C.c_int, [c_float_p] * 9 + [c_int_p, C.c_void_p] + [c_float_p] * 4 +
                                [C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, c_int_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nsac_score_pack_vecs_offset": (C.c_size_t, [C.c_int]),
    "nsac_score_aggregate_tc_cv": (C.c_int, [c_float_p] * 9 + [c_int_p, C.c_void_p, C.c_void_p] + [c_float_p] * 4 +
                                   [C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, c_int_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nsac_debug_score_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "nsac_debug_gemm_trace": (C.c_int, [C.c_void_p]),
    "nsac_camera_errors": (C.c_int, [c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_void_p]),
    "nsac_stem_im2col_planes": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsac_maxpool3x3s2_nhwc": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "nsac_subsample2_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "nsac_add_relu_nhwc": (C.c_int, [c_float_p, c_float_p, C.c_size_t, C.c_int, c_float_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsac_stem_im2col_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsac_stem_im2col_u8_cls": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsac_stem_border_fix": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), c_float_p, C.c_void_p]),
    "nsac_im2col3x3_from_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
    "nsac_plane_post_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "nsac_plane_postprocess": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p] + [C.c_int] * 7 +
                               [C.c_float, C.c_float, C.c_double] + [C.c_void_p] * 12),
    "nsac_prune_assignment": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, c_float_p, C.c_void_p]),
    "nsac_pose_canon": (C.c_int, [c_float_p, c_float_p, C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "nsac_cam_rows": (C.c_int, [c_float_p, c_float_p, C.c_int, c_float_p, C.c_void_p]),
    "nsac_pixel_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "nsac_pixel_forward": (C.c_int, [C.POINTER(PixelWeights)] + [C.c_void_p] * 6 + [C.c_int] * 3 + [c_float_p] * 8 +
                           [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p]),
    "nsac_head_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "nsac_head_forward": (C.c_int, [C.POINTER(HeadWeights)] + [C.c_void_p] * 6 + [C.c_int] * 3 + [c_float_p] * 4 + [C.c_void_p] * 2 +
                          [C.c_int] * 2 + [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 20 +
                          [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "nsac_model_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "nsac_model_forward": (C.c_int, [C.POINTER(BackboneWeights), C.POINTER(HeadWeights), C.c_void_p] + [C.c_int] * 3 + [c_float_p] * 4 +
                           [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 20 +
                           [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "nsac_backbone_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "nsac_backbone_forward": (C.c_int, [C.POINTER(BackboneWeights), C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 +
                              [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p]),
    "nsac_match_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "nsac_match_forward": (C.c_int, [C.POINTER(MatchWeights)] + [c_float_p] * 5 + [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                    C.c_int, c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p]),
    "nsac_refine_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "nsac_refine_forward": (C.c_int, [C.POINTER(RefineWeights), c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_int, c_float_p, c_float_p,
                                      c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 12 +
                            [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("NSAC_B200_LIB", LIB_PATH)      # development aid: A/B a kernel variant built elsewhere (scripts/)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nopesac_b200 has no CPU / eager fallback).")
        L = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().nsac_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")

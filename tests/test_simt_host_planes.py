"""CPU: the SOURCE of nopesac_b200/csrc/planes.cu executed on the host (tests/simt_host: one OS thread per CUDA thread) and
checked against the plane post-processing oracle — kernel logic is exercised here because the build container has no GPU; the
same comparison runs against the real device build in tests/test_gpu_planes.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_host"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import run as simt_run  # noqa: E402
from planes_check import check_image  # noqa: E402

from nopesac_b200 import synthetic  # noqa: E402
from oracle import planes_restate  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    L = simt_run.build("planes.cu")
    L.nsac_plane_post_workspace_bytes.restype = C.c_size_t
    L.nsac_plane_post_workspace_bytes.argtypes = [C.c_int] * 4
    L.nsac_plane_postprocess.restype = C.c_int
    L.nsac_plane_postprocess.argtypes = [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_float, C.c_float, C.c_double] + [C.c_void_p] * 12
    L.simt_last_error.restype = C.c_char_p
    return L


def _aligned(nbytes, align=256):
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


def host_postprocess(L, batch, H, W, thr=(0.6, 0.5, 0.6)):
    logits, params, masks, feats = (np.ascontiguousarray(batch[k].numpy()) for k in ("pred_logits", "pred_params", "pred_mask_logits", "query_feat"))
    B, NQ = logits.shape[:2]
    h, w = masks.shape[-2:]
    Cf = feats.shape[-1]
    ws = _aligned(L.nsac_plane_post_workspace_bytes(B, NQ, H, W))
    ws[:] = 0xA5                                             # the kernels must not rely on a zeroed workspace
    out = {"count": np.full(B, -7, np.int32), "flags": np.full(B, -7, np.int32), "ori_idx": np.full((B, NQ), -7, np.int32),
           "planes": np.full((B, NQ, 3), np.nan, np.float32), "feats": np.full((B, NQ, Cf), np.nan, np.float32),
           "scores": np.full((B, NQ), np.nan, np.float32), "centers": np.full((B, NQ, 2), np.nan, np.float32),
           "bboxes": np.full((B, NQ, 4), np.nan, np.float32), "areas": np.full((B, NQ), -7, np.int32)}
    seg = _aligned(B * H * W, 16)
    seg[:] = 0x77
    p = lambda a: C.c_void_p(a.ctypes.data)
    st = L.nsac_plane_postprocess(p(logits), p(params), p(masks), p(feats), B, NQ, Cf, h, w, H, W, thr[0], thr[1], thr[2],
                                  p(out["count"]), p(out["flags"]), p(out["ori_idx"]), p(out["planes"]), p(out["feats"]),
                                  p(out["scores"]), p(out["centers"]), p(out["bboxes"]), p(out["areas"]), p(seg), p(ws), None)
    assert st == 0, L.simt_last_error()
    out["seg"] = seg.reshape(B, H, W)
    return out


@pytest.mark.parametrize("nq,h,w,scale", [(12, 9, 13, 4), (50, 30, 40, 4), (20, 17, 50, 2), (127, 8, 8, 4)])
def test_planes_kernel_source_on_host_matches_oracle(lib, nq, h, w, scale):
    cases = synthetic.PLANE_HEAD_CASES
    B = len(cases)
    batch = synthetic.make_plane_head_batch(500 + nq, B, cases=cases, num_queries=nq, mask_h=h, mask_w=w, channels=8)
    H, W = h * scale, w * scale
    got = host_postprocess(lib, batch, H, W)
    want = planes_restate.postprocess_plane_head_mask(batch["pred_logits"], batch["pred_params"], batch["pred_mask_logits"],
                                                      batch["query_feat"], H, W)
    ties = 0
    for b in range(B):
        ties += check_image({k: v[b] for k, v in got.items()}, want[b], tag=f"{cases[b]} nq={nq} {h}x{w} x{scale}")
    assert ties <= 4


def test_planes_entry_point_rejects_bad_arguments(lib):
    batch = synthetic.make_plane_head_batch(0, 1, num_queries=8, mask_h=8, mask_w=8, channels=4)
    with pytest.raises(AssertionError, match="2x or 4x"):
        host_postprocess(lib, batch, 24, 24)
    big = synthetic.make_plane_head_batch(0, 1, num_queries=128, mask_h=8, mask_w=8, channels=4)
    with pytest.raises(AssertionError, match="bad shape"):
        host_postprocess(lib, big, 32, 32)

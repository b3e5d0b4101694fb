"""Shared comparison of a `nsac_plane_postprocess` result (numpy arrays of ONE image) with the oracle's result dict — used
by the GPU parity tests and by the host execution of the kernel source (tests/test_simt_host_planes.py).

Bar: bit-exact on the index path (kept list, query indices, gathered params / features, flags) and on the label map wherever
the oracle's decision is not a float near-tie.  `sigmoid` goes through `expf`, whose last bit differs between libraries
(ATen/Sleef on the host, CUDA, glibc), so a pixel may differ ONLY where the oracle's own margin — top-1 minus top-2 of
score * prob, or |top-1 - MASK_PROB_THRESHOLD| — is below `tie_tol`; such pixels are counted and bounded, areas may differ by at
most that count, centres / boxes are then compared with a matching tolerance.
"""
import numpy as np
import torch

FLAG_ZERO, FLAG_FALLBACK, FLAG_PATCH00 = 1, 2, 4


def check_image(got: dict, o: dict, mask_thr: float = 0.5, tie_tol: float = 2e-6, max_tie_frac: float = 2e-3, tag=""):
    n = int(got["count"])
    NQ = got["ori_idx"].shape[0]
    assert n == len(o["pred_plane_oriIdxs"]), (tag, n, o["pred_plane_oriIdxs"])
    assert got["ori_idx"][:n].tolist() == o["pred_plane_oriIdxs"], (tag, got["ori_idx"][:n], o["pred_plane_oriIdxs"])
    assert (got["ori_idx"][n:] == -1).all(), tag
    flags = int(got["flags"])
    assert bool(flags & FLAG_ZERO) == o["zero_flag"] and bool(flags & FLAG_FALLBACK) == o["fallback"], (tag, flags)
    assert np.array_equal(got["planes"][:n], o["pred_plane"].numpy()), tag                      # gathers: bit-exact
    assert np.array_equal(got["feats"][:n], o["pred_plane_feats"][0].numpy()), tag
    assert not got["planes"][n:].any() and not got["feats"][n:].any(), tag
    assert np.abs(got["scores"][:n] - np.asarray(o["scores"], dtype=np.float32)).max() <= 1e-6, tag

    H, W = got["seg"].shape
    want = np.full((H, W), 255, dtype=np.uint8)
    for j in range(n):
        want[o["pred_plane_masks"][j].numpy()] = j
    diff = got["seg"] != want
    nd = int(diff.sum())
    if nd:
        vw = o["_valid_w"]
        top = torch.topk(vw, k=min(2, vw.shape[0]), dim=0).values
        margin = (top[0] - top[1]) if vw.shape[0] > 1 else torch.full_like(top[0], float("inf"))
        near = ((margin < tie_tol) | ((top[0] - mask_thr).abs() < tie_tol)).numpy()
        assert near[diff].all(), (tag, "label map differs away from ties", nd, int((diff & ~near).sum()))
        assert nd <= max_tie_frac * H * W, (tag, nd)
    assert np.abs(got["areas"][:n].astype(np.int64) - np.asarray(o["areas"], dtype=np.int64)).max() <= nd, (tag, got["areas"][:n], o["areas"])
    want_c = o["pred_plane_ins_center"].numpy()
    want_b = np.asarray(o["bboxes"], dtype=np.float32).reshape(n, 4)
    if nd == 0:
        assert np.array_equal(got["bboxes"][:n], want_b), (tag, got["bboxes"][:n], want_b)
        assert np.allclose(got["centers"][:n], want_c, rtol=0, atol=1e-6, equal_nan=True), (tag, got["centers"][:n], want_c)
    else:
        assert np.abs(got["bboxes"][:n] - want_b).max() <= max(H, W), tag
        assert np.allclose(got["centers"][:n], want_c, rtol=0, atol=1e-3, equal_nan=True), (tag, got["centers"][:n], want_c)
    return nd

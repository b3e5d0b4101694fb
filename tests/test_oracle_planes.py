"""CPU: the plane post-processing oracle (oracle/planes_restate.py, row f1) against the reference's own method (when
/root/reference is present), against the committed golden fixture generated from it, and the host-side pieces of
nopesac_b200.plane_postprocess (RLE strings, refusal of CPU tensors)."""
import os

import numpy as np
import pytest
import torch

from nopesac_b200 import synthetic
from oracle import planes_restate, ref_planes_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rle_decode(counts: bytes, h: int, w: int) -> np.ndarray:
    """Inverse of rleToString + rleEncode (pycocotools rleFrString / rleDecode)."""
    runs, p, s = [], 0, counts
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(runs) > 2:
            x += runs[-2]
        runs.append(x)
    flat = np.zeros(h * w, dtype=bool)
    pos, v = 0, False
    for r in runs:
        flat[pos:pos + r] = v
        pos += r
        v = not v
    return flat.reshape((h, w), order="F")


def _oracle(it, h=480, w=640):
    return planes_restate.postprocess_plane_head_mask(it["pred_logits"][None], it["pred_params"][None], it["pred_mask_logits"][None],
                                                      it["query_feat"][None], h, w)[0]


def test_rle_round_trip_and_bbox():
    rng = np.random.default_rng(1)
    for t in range(60):
        h, w = rng.integers(1, 30, 2)
        m = rng.random((h, w)) < rng.random()
        if t % 9 == 0:
            m[:] = (t // 9) % 2
        counts = planes_restate.rle_encode(m)
        assert sum(counts) == h * w
        assert np.array_equal(rle_decode(planes_restate.rle_to_string(counts), h, w), m)
        bb = planes_restate.rle_to_bbox(counts, h, w)
        if m.any():
            ys, xs = np.nonzero(m)
            assert bb == [xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1]
        else:
            assert bb == [0, 0, 0, 0]


def test_planes_oracle_matches_golden():
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "planes_post.golden"), weights_only=False)
    H, W = fx["height"], fx["width"]
    seen = set()
    for rec in fx["records"]:
        it = synthetic.make_plane_head_outputs(rec["image_idx"], num_queries=fx["num_queries"], case=rec["case"])
        o = _oracle(it, H, W)
        seen.add((rec["case"], o["zero_flag"], o["fallback"]))
        assert o["pred_plane_oriIdxs"] == rec["pred_plane_oriIdxs"], rec["case"]
        assert torch.equal(o["pred_plane"], rec["pred_plane"]) and torch.equal(o["pred_plane_feats"], rec["pred_plane_feats"])
        assert np.array_equal(o["pred_plane_ins_center"].numpy(), rec["pred_plane_ins_center"].numpy(), equal_nan=True)
        assert o["scores"] == rec["scores"] and o["bboxes"] == rec["bboxes"] and o["counts"] == rec["counts"] and o["areas"] == rec["areas"]
        for j, c in enumerate(rec["counts"]):
            assert np.array_equal(rle_decode(c, H, W), o["pred_plane_masks"][j].numpy())
    # every branch of the reference's post-processing is in the fixture
    assert ("regular", False, False) in seen and ("zero", True, False) in seen and ("zero_empty", True, False) in seen
    assert ("fallback_empty", False, True) in seen and ("fallback_overlap", False, True) in seen


@pytest.mark.skipif(not ref_planes_loader.available(), reason="reference tree not present")
def test_planes_oracle_matches_reference_source():
    cases = synthetic.PLANE_HEAD_CASES
    for (nq, h, w, scale) in ((50, 120, 160, 4), (20, 30, 40, 4), (20, 60, 80, 2)):
        b = synthetic.make_plane_head_batch(900, len(cases), cases=cases, num_queries=nq, mask_h=h, mask_w=w, channels=32)
        H, W = h * scale, w * scale
        run = ref_planes_loader.load(num_queries=nq, height=H, width=W)
        ref = run({k: b[k] for k in ("pred_logits", "pred_params", "pred_mask_logits")}, b["query_feat"])
        got = planes_restate.postprocess_plane_head_mask(b["pred_logits"], b["pred_params"], b["pred_mask_logits"], b["query_feat"], H, W)
        for c, r, o in zip(cases, ref, got):
            assert torch.equal(r["pred_plane"], o["pred_plane"]), c
            assert torch.equal(r["pred_plane_feats"], o["pred_plane_feats"]), c
            assert [int(x) for x in r["pred_plane_oriIdxs"]] == o["pred_plane_oriIdxs"], c
            assert torch.equal(r["pred_plane_masks"].bool(), o["pred_plane_masks"]), c
            assert np.array_equal(r["pred_plane_ins_center"].numpy(), o["pred_plane_ins_center"].numpy(), equal_nan=True), c
            assert [ins["score"] for ins in r["instances"]] == o["scores"], c
            assert [ins["bbox"] for ins in r["instances"]] == o["bboxes"], c
            assert [ins["segmentation"]["counts"] for ins in r["instances"]] == o["counts"], c


def test_product_rle_matches_oracle_rle():
    from nopesac_b200.plane_postprocess import rle_counts
    rng = np.random.default_rng(0)
    for t in range(100):
        h, w = rng.integers(1, 40, 2)
        m = rng.random((h, w)) < rng.random()
        if t % 7 == 0:
            m[:] = t % 2
        assert rle_counts(m) == planes_restate.rle_to_string(planes_restate.rle_encode(m))


def test_plane_postprocess_refuses_cpu_tensors():
    from nopesac_b200 import plane_postprocess
    it = synthetic.make_plane_head_outputs(0, num_queries=8, mask_h=8, mask_w=8, channels=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        plane_postprocess.postprocess_plane_head_mask({k: it[k][None] for k in ("pred_logits", "pred_params", "pred_mask_logits")},
                                                      it["query_feat"][None], 32, 32)


def test_plane_workspace_size_is_host_only():
    from nopesac_b200 import _lib
    L = _lib.lib()
    assert L.nsac_plane_post_workspace_bytes(0, 50, 480, 640) == 0
    one, two = L.nsac_plane_post_workspace_bytes(1, 50, 480, 640), L.nsac_plane_post_workspace_bytes(2, 50, 480, 640)
    assert one >= 480 * 640 and one % 256 == 0 and two > one

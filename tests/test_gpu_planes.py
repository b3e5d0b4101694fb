"""GPU: `nsac_plane_postprocess` / nopesac_b200.plane_postprocess (row f1: the plane lists that feed the camera head) against
the oracle (oracle/planes_restate.py, pinned to the reference's `_postprocess_planeHeadMask`) and the golden fixture generated
from the reference.  Bar (tests/planes_check.py): bit-exact kept lists, indices, gathers and flags; label map identical except
at float near-ties of the oracle's own margins (expf differs in the last bit between libraries), which are counted."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from planes_check import check_image  # noqa: E402

KEYS = ("count", "flags", "ori_idx", "planes", "feats", "scores", "centers", "bboxes", "areas", "seg")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _run(batch, H, W, dev):
    from nopesac_b200 import plane_postprocess
    outs = {k: batch[k].to(dev) for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    res = plane_postprocess.postprocess_plane_head_mask(outs, batch["query_feat"].to(dev), H, W)
    return res, {k: getattr(res, k).cpu().numpy() for k in KEYS}


def _oracle(batch, H, W):
    from oracle import planes_restate
    return planes_restate.postprocess_plane_head_mask(batch["pred_logits"], batch["pred_params"], batch["pred_mask_logits"],
                                                      batch["query_feat"], H, W)


@pytest.mark.parametrize("nq,h,w,scale", [(12, 9, 13, 4), (50, 30, 40, 4), (20, 17, 50, 2), (127, 8, 8, 4), (50, 120, 160, 4)])
def test_plane_postprocess_matches_oracle(nq, h, w, scale):
    dev = _dev()
    from nopesac_b200 import synthetic
    cases = synthetic.PLANE_HEAD_CASES
    B = len(cases)
    batch = synthetic.make_plane_head_batch(500 + nq, B, cases=cases, num_queries=nq, mask_h=h, mask_w=w, channels=8 if h < 120 else 256)
    H, W = h * scale, w * scale
    _, got = _run(batch, H, W, dev)
    want = _oracle(batch, H, W)
    ties = sum(check_image({k: v[b] for k, v in got.items()}, want[b], tag=f"{cases[b]} nq={nq} {h}x{w} x{scale}") for b in range(B))
    assert ties <= 1e-4 * B * H * W


def test_plane_postprocess_matches_reference_golden():
    dev = _dev()
    from nopesac_b200 import synthetic
    from test_oracle_planes import rle_decode
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "planes_post.golden"), weights_only=False)
    H, W, NQ = fx["height"], fx["width"], fx["num_queries"]
    items = [synthetic.make_plane_head_outputs(r["image_idx"], num_queries=NQ, case=r["case"]) for r in fx["records"]]
    batch = {k: torch.stack([it[k] for it in items]).contiguous() for k in items[0]}
    res, got = _run(batch, H, W, dev)
    want = _oracle(batch, H, W)
    for b, rec in enumerate(fx["records"]):
        n = int(got["count"][b])
        assert got["ori_idx"][b, :n].tolist() == rec["pred_plane_oriIdxs"], rec["case"]
        assert np.array_equal(got["planes"][b, :n], rec["pred_plane"].numpy())
        assert np.array_equal(got["feats"][b, :n], rec["pred_plane_feats"][0].numpy())
        nd = check_image({k: v[b] for k, v in got.items()}, want[b], tag=rec["case"])
        if nd == 0:                                        # then the masks are the reference's, bit for bit
            for j, c in enumerate(rec["counts"]):
                assert np.array_equal(got["seg"][b] == j, rle_decode(c, H, W)), (rec["case"], j)
            assert got["areas"][b, :n].tolist() == rec["areas"]
            assert got["bboxes"][b, :n].tolist() == rec["bboxes"]
    # a second call on the same inputs gives identical bytes (integer / fixed-point accumulation: no atomic-order effects)
    _, again = _run(batch, H, W, dev)
    for k in KEYS:
        assert np.array_equal(got[k], again[k], equal_nan=(got[k].dtype.kind == "f")), k


def test_reference_result_format_and_meta_arch_method():
    dev = _dev()
    from nopesac_b200 import config, meta_arch, synthetic
    batch = synthetic.make_plane_head_batch(300, 2, cases=("regular", "zero"))
    want = _oracle(batch, 480, 640)
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg()).to(dev)
    outs = {k: batch[k].to(dev) for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    metas = [{"image_id": f"img{i}", "file_name": f"img{i}.png", "height": 480, "width": 640} for i in range(2)]
    res = model._postprocess_planeHeadMask(outs, [None, None], metas, [(480, 640)] * 2, batch["query_feat"].to(dev))
    assert len(res) == 2
    for r, o, m in zip(res, want, metas):
        n = len(o["pred_plane_oriIdxs"])
        assert r["image_id"] == m["image_id"] and r["file_name"] == m["file_name"]
        assert r["pred_plane"].shape == (n, 3) and r["pred_plane_feats"].shape == (1, n, 256)
        assert r["pred_plane_masks"].shape == (n, 480, 640) and r["pred_plane_masks"].dtype == torch.bool
        assert r["pred_plane_ins_center"].shape == (n, 2)
        assert [int(x) for x in r["pred_plane_oriIdxs"]] == o["pred_plane_oriIdxs"]
        assert torch.equal(r["pred_plane"].cpu(), o["pred_plane"]) and torch.equal(r["pred_plane_feats"].cpu(), o["pred_plane_feats"])
        assert len(r["instances"]) == n
        for j, ins in enumerate(r["instances"]):
            assert ins["category_id"] == 0 and ins["bbox_mode"] == 1 and ins["segmentation"]["size"] == [480, 640]
            assert abs(ins["score"] - o["scores"][j]) <= 1e-6
            if torch.equal(r["pred_plane_masks"][j].cpu(), o["pred_plane_masks"][j]):
                assert ins["bbox"] == o["bboxes"][j] and ins["segmentation"]["counts"] == o["counts"][j]


def test_plane_postprocess_fails_loudly():
    dev = _dev()
    from nopesac_b200 import plane_postprocess, synthetic
    small = synthetic.make_plane_head_batch(0, 1, num_queries=8, mask_h=8, mask_w=8, channels=4)
    outs = {k: small[k].to(dev) for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    with pytest.raises(RuntimeError, match="2x or 4x"):
        plane_postprocess.postprocess_plane_head_mask(outs, small["query_feat"].to(dev), 24, 24)
    big = synthetic.make_plane_head_batch(0, 1, num_queries=128, mask_h=8, mask_w=8, channels=4)
    outs = {k: big[k].to(dev) for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    with pytest.raises(RuntimeError, match="bad shape"):
        plane_postprocess.postprocess_plane_head_mask(outs, big["query_feat"].to(dev), 32, 32)

"""Generates tests/golden/planes_post.golden from the LIVE reference (`_postprocess_planeHeadMask` cut out of
/root/reference/NopeSAC_Net/modeling/meta_arch/siamese_planeTR.py by oracle/ref_planes_loader.py) on the seeded synthetic
PlaneTRHead outputs of nopesac_b200.synthetic.make_plane_head_outputs.  Inputs are regenerated from (image index, case) by the
tests; the fixture stores only the reference's results (masks as COCO RLE strings, so the file stays small).

    python tests/golden/make_planes_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nopesac_b200 import synthetic  # noqa: E402
from oracle import ref_planes_loader  # noqa: E402

CASES = [(300 + i, c) for i, c in enumerate(synthetic.PLANE_HEAD_CASES * 2)]


def main():
    run = ref_planes_loader.load(num_queries=50)
    records = []
    for idx, case in CASES:
        it = synthetic.make_plane_head_outputs(idx, case=case)
        outs = {k: it[k][None] for k in ("pred_logits", "pred_params", "pred_mask_logits")}
        r = run(outs, it["query_feat"][None])[0]
        records.append({
            "image_idx": idx, "case": case,
            "pred_plane": r["pred_plane"].clone(),
            "pred_plane_feats": r["pred_plane_feats"].clone(),
            "pred_plane_oriIdxs": [int(x) for x in r["pred_plane_oriIdxs"]],
            "pred_plane_ins_center": r["pred_plane_ins_center"].clone(),
            "scores": [ins["score"] for ins in r["instances"]],
            "bboxes": [ins["bbox"] for ins in r["instances"]],
            "counts": [ins["segmentation"]["counts"] for ins in r["instances"]],
            "areas": [int(m.sum()) for m in r["pred_plane_masks"]],
        })
    out = os.path.join(ROOT, "tests", "golden", "planes_post.golden")
    torch.save({"height": 480, "width": 640, "num_queries": 50, "records": records}, out)
    print(out, os.path.getsize(out), "bytes;", [(r["case"], len(r["areas"])) for r in records])


if __name__ == "__main__":
    main()

"""Times nsac_plane_postprocess (row f1) on the device: B images of the inference_mp3d shape (50 queries, 120x160 mask logits ->
480x640), CUDA events on the launch stream, inputs > L2 for B >= 64.  Prints ms per call, images/s and the algorithmic GB/s
(4 NQ h w + 3 H W bytes per image).   python scripts/planes_bench.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nopesac_b200 import plane_postprocess, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128]
    base = synthetic.make_plane_head_batch(300, 8, cases=("regular",))
    for B in sizes:
        rep = (B + 7) // 8
        batch = {k: v.repeat(rep, *([1] * (v.dim() - 1)))[:B].contiguous().to(dev) for k, v in base.items()}
        outs = {k: batch[k] for k in ("pred_logits", "pred_params", "pred_mask_logits")}
        for _ in range(3):
            res = plane_postprocess.postprocess_plane_head_mask(outs, batch["query_feat"], 480, 640)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            res = plane_postprocess.postprocess_plane_head_mask(outs, batch["query_feat"], 480, 640)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        nq, h, w = batch["pred_mask_logits"].shape[1:]
        alg = B * (4 * nq * h * w + 3 * 480 * 640)
        print(f"B={B} planes/img={res.count.float().mean().item():.1f} {ms:.3f} ms/call {B / ms * 1e3:.0f} images/s "
              f"{alg / ms / 1e6:.1f} GB/s algorithmic ({alg / 1e6:.1f} MB)", flush=True)


if __name__ == "__main__":
    main()

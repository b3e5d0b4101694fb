"""One launch family of the tcgen05 GEMM engine per shape — the target of `ncu --set full -k regex:gemm_bf16x3` captures and of
the tensor-pipe roofline entry (timed with CUDA events when run without ncu).
    python scripts/profile_gemm.py k7|conv256|gnn|all [reps]
  k7       K7 hypothesis-generation layer: 16384 x 1024 x 1024 (B*NQ rows of configs[1])
  conv256  3x3 convolution 256 -> 256 on the 60 x 80 level, 128 images (convs_backbone.0 of the pixel pose network)
  gnn      one GNN linear: 2048 x 256 x 256 (64 pairs x 2 views x 16 planes)
  res      residual-epilogue GEMM of the backbone: res2.x.conv3, 128 images x 120 x 160 pixels, 64 -> 256 + shortcut + ReLU (HBM-bound:
           reports algorithmic GB/s = A + residual + output planes)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)


def planes(rows, K, scale=1.0):
    return ops.split(rnd(rows, K) * scale)


def timed(fn, flops, name, passes=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"shape": name, "us": ms * 1e3, "tflops_fp32_equiv": flops / ms / 1e9, "tflops_tensor_fp16": passes * flops / ms / 1e9,
                      "passes": passes}), flush=True)


if which in ("k7", "all"):
    M, N, K = 16384, 1024, 1024
    a, w = planes(M, K), planes(N, K, 0.03)
    out = ops.Split.empty(M, N, dev)
    timed(lambda: ops.gemm_tc(a, w, None, ops.ACT_RELU, want_f32=False, out_split=out), 2.0 * M * N * K, f"k7 {M}x{N}x{K}")
if which in ("conv256", "all"):
    Nimg, H, W, C = 128, 60, 80, 256
    x = planes(Nimg * H * W, C)
    w = planes(C, 9 * C, 0.02)
    timed(lambda: ops.conv3x3_tc(x, Nimg, H, W, w, None, ops.ACT_LEAKY, want_f32=False, want_split=True),
          2.0 * Nimg * H * W * C * 9 * C, f"conv3x3 {Nimg}x{H}x{W} {C}->{C}")
if which in ("gnn", "all"):
    M, N, K = 2048, 256, 256
    a, w = planes(M, K), planes(N, K, 0.06)
    timed(lambda: ops.gemm_tc(a, w), 2.0 * M * N * K, f"gnn {M}x{N}x{K}")
if which in ("res", "all"):
    M, N, K = 128 * 120 * 160, 256, 64
    a, w = planes(M, K), planes(N, K, 0.1)
    res = ops.Split.empty(M, N, dev)
    res.hi.normal_(); res.lo.zero_()
    out = ops.Split.empty(M, N, dev)
    bias = rnd(N)
    fn = lambda: ops.gemm_tc(a, w, bias, ops.ACT_RELU, want_f32=False, out_split=out, residual=res)
    timed(fn, 2.0 * M * N * K, f"res2.conv3 {M}x{N}x{K} + residual")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = M * (K * 4 + N * 4 + N * 4)
    print(json.dumps({"shape": "res2.conv3 residual epilogue", "us": ms * 1e3, "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6}), flush=True)
torch.cuda.synchronize()

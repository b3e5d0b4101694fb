"""Micro-benchmark of the tcgen05 split-bf16 GEMM engine (CUDA events, inputs > L2 not required: operands
are L2-resident by design in the path).  Prints effective fp32 TFLOP/s (2MNK / t) and bf16-pass TFLOP/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops

dev = torch.device("cuda:0")
shapes = [(16384, 1024, 1024), (16384, 1024, 1280), (16384, 512, 1024), (16384, 512, 512), (16384, 256, 512),
          (2048, 768, 256), (614400, 256, 2304), (614400, 128, 1152)]
for (M, N, K) in shapes:
    for passes in (3, 1):
        a = ops.Split(torch.randn(M, K, device=dev).half(), torch.randn(M, K, device=dev).half() * 0.01, K)
        w = ops.Split(torch.randn(N, K, device=dev).half(), torch.randn(N, K, device=dev).half() * 0.01, K)
        out = ops.Split.empty(M, N, dev)
        for _ in range(3):
            ops.gemm_tc(a, w, None, 1, passes=passes, want_f32=False, out_split=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.gemm_tc(a, w, None, 1, passes=passes, want_f32=False, out_split=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * M * N * K
        print(f"M={M} N={N} K={K} passes={passes}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s fp32-equivalent  "
              f"{passes*fl/ms/1e9:7.1f} TFLOP/s bf16 tensor")
        del a, w, out

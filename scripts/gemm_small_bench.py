"""Fixed cost of one launch of the tcgen05 GEMM engine: chains of 20 identical GEMMs captured in a CUDA graph and
replayed (per-GEMM device time, no host overhead)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops

dev = torch.device("cuda:0")
shapes = [(128, 128, 64), (128, 128, 256), (2048, 256, 64), (2048, 256, 256), (2048, 512, 512), (2048, 256, 512), (2048, 768, 256),
          (16384, 256, 512), (16384, 1024, 1024)]
for (M, N, K) in shapes:
    for passes in (3, 1):
        a = ops.Split(torch.randn(M, K, device=dev).half(), torch.randn(M, K, device=dev).half() * 0.01, K)
        w = ops.Split(torch.randn(N, K, device=dev).half(), torch.randn(N, K, device=dev).half() * 0.01, K)
        out = ops.Split.empty(M, N, dev)
        run = lambda: ops.gemm_tc(a, w, None, 1, passes=passes, want_f32=False, out_split=out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 100 * 1e3
        fl = 2.0 * M * N * K
        print(f"M={M:6d} N={N:5d} K={K:5d} passes={passes}: {us:8.2f} us per GEMM   {passes*fl/us/1e6:7.1f} TFLOP/s fp16 tensor", flush=True)

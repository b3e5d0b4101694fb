"""Timeline of one small GEMM-engine launch (CTA 0): where do the ~12 us of a single-tile launch go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nopesac_b200 import ops, _lib

dev = torch.device("cuda:0")
names = ["entry", "setup done", "first TMA issued", "first operands landed", "first chunk committed", "epilogue start", "epilogue end",
         "after final barrier", "after dealloc", "accumulators read"]
for (M, N, K) in [(128, 128, 64), (2048, 256, 256), (2048, 512, 512)]:
    a = ops.Split(torch.randn(M, K, device=dev).half(), torch.randn(M, K, device=dev).half() * 0.01, K)
    w = ops.Split(torch.randn(N, K, device=dev).half(), torch.randn(N, K, device=dev).half() * 0.01, K)
    out = ops.Split.empty(M, N, dev)
    for _ in range(3):
        ops.gemm_tc(a, w, None, 1, passes=3, want_f32=False, out_split=out)
    torch.cuda.synchronize()
    bufs = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(3)]
    for b in bufs:          # three back-to-back launches: the gap between "after dealloc" and the next "entry" is the launch gap
        _lib.lib().nsac_debug_gemm_trace(b.data_ptr())
        ops.gemm_tc(a, w, None, 1, passes=3, want_f32=False, out_split=out)
    _lib.lib().nsac_debug_gemm_trace(None)
    torch.cuda.synchronize()
    t0 = int(bufs[0][0])
    print(f"M={M} N={N} K={K}")
    for i, b in enumerate(bufs):
        t = [int(x) - t0 for x in b[:10].tolist()]
        print(f"  launch {i}: " + "  ".join(f"{n} {v}" for n, v in zip(names, t)))

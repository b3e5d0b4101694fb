"""GPU: the tcgen05 split-precision GEMM engine (nsac_gemm_split / nsac_split16; fp16 or bf16 hi/lo planes,
1..4 MMA passes) against torch fp64."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rand(g, *shape):
    return torch.randn(*shape, generator=g)


def test_split_roundtrip():
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = _rand(g, 77, 100).to(dev)
    for fmt, dt, bits in ((ops.SPLIT_F16, torch.float16, 21), (ops.SPLIT_BF16, torch.bfloat16, 16)):
        s = ops.split(x, fmt)
        assert s.hi.shape == (77, 128) and s.hi.dtype == dt and s.lo.dtype == dt
        assert float(s.hi[:, 100:].float().abs().max()) == 0.0 and float(s.lo[:, 100:].float().abs().max()) == 0.0
        assert torch.equal(s.hi[:, :100], x.to(dt))
        assert util.maxdiff(s.float(), x) <= 2 ** -bits * float(x.abs().max())
    w = ops.split_weight(x * 1e-3)
    assert w.scale > 1.0 and util.maxdiff(w.float(), x * 1e-3) <= 2 ** -21 * 1e-3 * float(x.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 512), (1000, 768, 256), (257, 1024, 1280), (64, 512, 192),
                                   (4096, 1024, 1024), (130, 144, 128)])
def test_gemm_split_matches_fp64(M, N, K):
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = _rand(g, M, K), _rand(g, N, K) / K ** 0.5, _rand(g, N)
    ref = x.double() @ w.double().T + b.double()
    scale = float(ref.abs().max())
    errs = {}
    for fmt in (ops.SPLIT_F16, ops.SPLIT_BF16):
        xs, ws = ops.split(x.to(dev), fmt), ops.split_weight(w.to(dev), fmt)
        for passes in (1, 2, 3, 4):
            out, sp = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_NONE, passes=passes, want_f32=True, want_split=True)
            torch.cuda.synchronize()
            errs[(fmt, passes)] = util.maxdiff(out, ref) / scale
            assert util.maxdiff(sp.float(), out) <= (2 ** -20 if fmt == ops.SPLIT_F16 else 2 ** -15) * scale
            assert sp.hi.shape[1] % 64 == 0
            if sp.hi.shape[1] > N:
                assert float(sp.hi[:, N:].float().abs().max()) == 0.0
    print(f"M={M} N={N} K={K}: rel err fp16 planes passes1-4 = " + " ".join(f"{errs[(0, p)]:.1e}" for p in (1, 2, 3, 4)) +
          " | bf16 planes = " + " ".join(f"{errs[(1, p)]:.1e}" for p in (1, 2, 3, 4)))
    assert errs[(ops.SPLIT_F16, 3)] <= 2e-6, errs
    assert errs[(ops.SPLIT_BF16, 3)] <= 3e-5, errs
    # single-pass modes are diagnostics only (the product runs 3 passes): loose sanity bounds
    assert errs[(ops.SPLIT_F16, 1)] < 6e-3 and errs[(ops.SPLIT_BF16, 1)] < 2e-2
    # activations fused in the epilogue
    xs, ws = ops.split(x.to(dev)), ops.split_weight(w.to(dev))
    o1, _ = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_RELU)
    o2, _ = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_LEAKY)
    assert util.maxdiff(o1, torch.relu(ref)) <= 2e-6 * scale
    assert util.maxdiff(o2, torch.nn.functional.leaky_relu(ref, 0.01)) <= 2e-6 * scale


def test_gemm_split_strided_outputs_and_grouped_bias():
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(9)
    M, N, K, G = 512, 256, 256, 64
    x, w = _rand(g, M, K), _rand(g, N, K) / K ** 0.5
    gb = _rand(g, M // G, N)
    ref = x.double() @ w.double().T + gb.double().repeat_interleave(G, 0)
    wide = torch.zeros(M, 1280, device=dev)
    wide_split = ops.Split.empty(M, 1280, dev)
    wide_split.hi.zero_(); wide_split.lo.zero_()
    ops.gemm_tc(ops.split(x.to(dev)), ops.split_weight(w.to(dev)), gb.to(dev), ops.ACT_NONE, bias_group_rows=G,
                out_f32=wide[:, 1024:], want_split=True, out_split=wide_split.cols(1024, 1280))
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(wide[:, 1024:], ref) <= 2e-6 * scale
    assert float(wide[:, :1024].abs().max()) == 0.0
    assert util.maxdiff(wide_split.cols(1024, 1280).float(), ref) <= 4e-6 * scale
    assert float(wide_split.hi[:, :1024].float().abs().max()) == 0.0
    # the column-sliced planes feed straight back in as the K = 1280 operand of the next layer
    w2 = _rand(g, 128, 1280) / 1280 ** 0.5
    out, _ = ops.gemm_tc(wide_split, ops.split_weight(w2.to(dev)))
    full = torch.zeros(M, 1280, dtype=torch.float64); full[:, 1024:] = wide[:, 1024:].double().cpu()
    assert util.maxdiff(out, full @ w2.double().T) <= 4e-6 * scale


def test_fp16_plane_overflow_is_loud():
    """|x| > 65504 cannot be carried by fp16 planes: the result must be inf/NaN, never a finite wrong number."""
    dev = _dev()
    from nopesac_b200 import ops
    x = torch.full((128, 64), 1.0, device=dev); x[5, 3] = 1e6
    w = torch.eye(64, device=dev)
    out, _ = ops.gemm_tc(ops.split(x), ops.split_weight(w))
    assert not bool(torch.isfinite(out[5, 3]))
    assert bool(torch.isfinite(out[6]).all())


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 512, 128), (130, 64, 576), (4096, 1024, 256), (300, 200, 192), (77, 48, 64),
                                   (40000, 256, 64), (20000, 512, 128), (30000, 1024, 256), (5000, 2048, 512), (333, 200, 640)])    # many tiles per CTA: every tile buffer is reused; K >= 512: 128-wide tiles, residual from global
def test_gemm_residual_epilogue_matches_fp64(M, N, K):
    """nsac_gemm_split_residual: relu(x.W^T + b + residual) with the residual given as hi/lo planes (TMA-prefetched tile) — the
    fused `out += shortcut; relu` of the backbone's bottleneck blocks; full tiles, ragged M / N, and the 64-wide tile variant."""
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(3 * M + N + K)
    x, w, b, r = _rand(g, M, K), _rand(g, N, K) / K ** 0.5, _rand(g, N), _rand(g, M, N)
    xs, ws, rs = ops.split(x.to(dev)), ops.split_weight(w.to(dev)), ops.split(r.to(dev))
    ref = torch.relu(x.double() @ w.double().T + b.double() + rs.float().double().cpu())      # the planes ARE the residual (22 bits)
    _, sp = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_RELU, want_f32=False, want_split=True, residual=rs)      # planes out only
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(sp.float(), ref) <= 3e-6 * scale, util.maxdiff(sp.float(), ref) / scale
    if sp.hi.shape[1] > N:
        assert float(sp.hi[:, N:].float().abs().max()) == 0.0          # the zero padding of the output planes is untouched
    # residual planes that are a column slice of a wider buffer (row stride > N), output into a column slice as well
    wide = ops.split(_rand(g, M, N + 64).to(dev))
    rv = wide.cols(64, 64 + N)
    dst = ops.Split.empty(M, N + 128, dev)
    dst.hi.zero_(); dst.lo.zero_()
    ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_NONE, want_f32=False, out_split=dst.cols(64, 64 + N), residual=rv)
    ref2 = x.double() @ w.double().T + b.double() + rv.float().double().cpu()
    assert util.maxdiff(dst.cols(64, 64 + N).float(), ref2) <= 3e-6 * float(ref2.abs().max())
    assert float(dst.hi[:, :64].float().abs().max()) == 0.0 and float(dst.hi[:, 64 + N:].float().abs().max()) == 0.0
    with pytest.raises(RuntimeError):          # fp32 output is not part of the residual epilogue's contract
        ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_RELU, want_f32=True, residual=rs)


@pytest.mark.parametrize("M,N,K", [(1000, 64, 192), (4096, 64, 256), (130, 40, 576)])
def test_gemm_narrow_tiles_and_single_plane_a(M, N, K):
    """N <= 64 runs the 64-wide tile variant; `a.lo is None` (A exact in one fp16 plane, e.g. raw 8-bit pixels) skips the lo.hi pass."""
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    xi = torch.randint(0, 256, (M, K), generator=g).float()             # exact in fp16
    w, b = _rand(g, N, K) / K ** 0.5, _rand(g, N)
    ref = torch.relu(xi.double() @ w.double().T + b.double())
    a = ops.Split(xi.to(dev).half().contiguous(), None, K)
    ws = ops.split_weight(w.to(dev))
    out, sp = ops.gemm_tc(a, ws, b.to(dev), ops.ACT_RELU, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(out, ref) <= 2e-6 * scale, util.maxdiff(out, ref) / scale
    assert util.maxdiff(sp.float(), out) <= 2 ** -20 * scale
    x = _rand(g, M, K)
    out3, _ = ops.gemm_tc(ops.split(x.to(dev)), ws, b.to(dev))
    ref3 = x.double() @ w.double().T + b.double()
    assert util.maxdiff(out3, ref3) <= 2e-6 * float(ref3.abs().max())

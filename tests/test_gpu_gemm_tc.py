"""GPU: the tcgen05 split-bf16 GEMM engine (nsac_gemm_bf16x3 / nsac_split_bf16) against torch fp64."""
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
    s = ops.split(x)
    assert s.hi.shape == (77, 128) and s.hi.dtype == torch.bfloat16
    assert float(s.hi[:, 100:].float().abs().max()) == 0.0 and float(s.lo[:, 100:].float().abs().max()) == 0.0
    assert torch.equal(s.hi[:, :100], x.to(torch.bfloat16))
    assert util.maxdiff(s.float(), x) <= 2 ** -16 * float(x.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 512), (1000, 768, 256), (257, 1024, 1280), (64, 512, 192),
                                   (4096, 1024, 1024), (130, 144, 128)])
def test_gemm_bf16x3_matches_fp64(M, N, K):
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = _rand(g, M, K), _rand(g, N, K) / K ** 0.5, _rand(g, N)
    ref = x.double() @ w.double().T + b.double()
    xs, ws = ops.split(x.to(dev)), ops.split(w.to(dev))
    out, sp = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_NONE, passes=3, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(out, ref) <= 3e-5 * scale, f"3-pass error {util.maxdiff(out, ref) / scale:.2e}"
    assert util.maxdiff(sp.float(), out) <= 2 ** -15 * scale          # re-split planes carry the same values
    assert sp.hi.shape[1] % 64 == 0
    if sp.hi.shape[1] > N:
        assert float(sp.hi[:, N:].float().abs().max()) == 0.0
    out1, _ = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_RELU, passes=1)
    e1 = util.maxdiff(out1, torch.relu(ref)) / scale
    assert 1e-5 < e1 < 2e-2, f"1-pass (plain bf16) error {e1:.2e} out of the expected band"
    out2, _ = ops.gemm_tc(xs, ws, b.to(dev), ops.ACT_LEAKY, passes=2)
    e2 = util.maxdiff(out2, torch.nn.functional.leaky_relu(ref, 0.01)) / scale
    assert e2 < e1


def test_gemm_bf16x3_strided_outputs_and_grouped_bias():
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
    ops.gemm_tc(ops.split(x.to(dev)), ops.split(w.to(dev)), gb.to(dev), ops.ACT_NONE, bias_group_rows=G,
                out_f32=wide[:, 1024:], want_split=True, out_split=wide_split.cols(1024, 1280))
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(wide[:, 1024:], ref) <= 3e-5 * scale
    assert float(wide[:, :1024].abs().max()) == 0.0
    assert util.maxdiff(wide_split.cols(1024, 1280).float(), ref) <= 1e-4 * scale
    assert float(wide_split.hi[:, :1024].float().abs().max()) == 0.0
    # the column-sliced planes feed straight back in as the K = 1280 operand of the next layer
    w2 = _rand(g, 128, 1280) / 1280 ** 0.5
    out, _ = ops.gemm_tc(wide_split, ops.split(w2.to(dev)))
    full = torch.zeros(M, 1280, dtype=torch.float64); full[:, 1024:] = wide[:, 1024:].double().cpu()
    assert util.maxdiff(out, full @ w2.double().T) <= 1e-4 * scale

"""GPU: the pixel pose-network kernels (implicit-GEMM 3x3 convolution on the tensor-core engine, GroupNorm with
upsample-add, max-pool, correlation softmax, im2col, NCHW->NHWC planes) against PyTorch fp64 references."""
import pytest
import torch
import torch.nn.functional as F

from tests import util

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _nhwc(x):   # [N,C,H,W] -> [N*H*W, C]
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(N * H * W, C).contiguous()


@pytest.mark.parametrize("N,C,H,W,Cout", [(3, 64, 15, 20, 128), (2, 128, 30, 40, 256), (2, 256, 60, 80, 128), (1, 320, 15, 20, 256),
                                          (2, 128, 7, 9, 64)])
def test_conv3x3_implicit_gemm(N, C, H, W, Cout):
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(N * C + H)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / (9 * C) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), padding=1), 0.01)
    xp = ops.nchw_to_planes(x.to(dev))
    assert util.maxdiff(xp.float(), _nhwc(x)) <= 2 ** -20 * float(x.abs().max())
    wp = ops.split_weight(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dev))
    out, sp = ops.conv3x3_tc(xp, N, H, W, wp, b.to(dev), ops.ACT_LEAKY, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert util.maxdiff(out, _nhwc(ref)) <= 3e-6 * scale, util.maxdiff(out, _nhwc(ref)) / scale
    assert util.maxdiff(sp.float(), out) <= 2 ** -19 * scale


@pytest.mark.parametrize("N,C,H,W,Cout", [(2, 128, 60, 80, 128), (3, 64, 31, 45, 64), (1, 256, 30, 40, 256), (2, 512, 15, 20, 512), (2, 64, 8, 8, 96)])
def test_conv3x3_stride2_implicit_gemm(N, C, H, W, Cout):
    """nsac_conv3x3_split_strided (stride 2: the TMA tensor map's traversal stride picks every second pixel) == F.conv2d(stride=2,
    padding=1), even and odd map sizes; and == the explicit planes im2col + GEMM route it replaces."""
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(7 * N + C + H)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / (9 * C) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1, stride=2))
    xp = ops.nchw_to_planes(x.to(dev))
    wp = ops.split_weight(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dev))
    out, sp = ops.conv3x3_tc(xp, N, H, W, wp, b.to(dev), ops.ACT_RELU, want_f32=True, want_split=True, stride=2)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert out.shape[0] == N * ref.shape[2] * ref.shape[3]
    assert util.maxdiff(out, _nhwc(ref)) <= 3e-6 * scale, util.maxdiff(out, _nhwc(ref)) / scale
    assert util.maxdiff(sp.float(), out) <= 2 ** -19 * scale
    cols, Ho, Wo = ops.im2col3x3_from_planes(xp, N, H, W, 2)
    assert (Ho, Wo) == tuple(ref.shape[2:])
    out2, _ = ops.gemm_tc(cols, wp, b.to(dev), ops.ACT_RELU)
    assert util.maxdiff(out2, out) <= 3e-6 * scale


@pytest.mark.parametrize("N,C,H,W,Cout", [(2, 256, 120, 160, 512), (3, 64, 31, 45, 64), (2, 1024, 30, 40, 2048), (1, 128, 9, 7, 96)])
def test_conv1x1_stride2_implicit_gemm(N, C, H, W, Cout):
    """nsac_conv1x1_split_strided (the strided projection shortcut of res3.0 / res4.0 / res5.0, gathered through the tensor map's
    traversal stride) == F.conv2d(kernel 1, stride 2, no padding), even and odd map sizes; == subsample + plain GEMM bit for bit."""
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(11 * N + C + H)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(Cout, C, generator=g) / C ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.double(), w.double()[:, :, None, None], b.double(), stride=2)
    xp = ops.nchw_to_planes(x.to(dev))
    wp = ops.split_weight(w.contiguous().to(dev))
    sp = ops.conv1x1_tc_strided(xp, N, H, W, wp, b.to(dev), ops.ACT_NONE, stride=2)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert sp.rows == N * ref.shape[2] * ref.shape[3]
    assert util.maxdiff(sp.float(), _nhwc(ref)) <= 2 ** -18 * scale, util.maxdiff(sp.float(), _nhwc(ref)) / scale
    sub = ops.subsample2_planes(xp, N, H, W)[0]
    _, sp2 = ops.gemm_tc(sub, wp, b.to(dev), ops.ACT_NONE, want_f32=False, want_split=True)
    assert torch.equal(sp.hi, sp2.hi) and torch.equal(sp.lo, sp2.lo)


def test_groupnorm_upsample_add():
    dev = _dev()
    from nopesac_b200 import ops
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 3, 128, 30, 40
    x = torch.randn(N, C, H, W, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    skip = torch.randn(N, C, H // 2, W // 2, generator=g)
    ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5)
    ref_relu = F.relu(ref)
    ref_skip = ref + F.interpolate(skip.double(), size=(H, W), mode="nearest")
    o1, _ = ops.groupnorm_nhwc(_nhwc(x).to(dev), N, H, W, gamma.to(dev), beta.to(dev), 32, 1e-5, True)
    o2, s2 = ops.groupnorm_nhwc(_nhwc(x).to(dev), N, H, W, gamma.to(dev), beta.to(dev), 32, 1e-5, False, _nhwc(skip).to(dev),
                                want_f32=True, want_split=True)
    torch.cuda.synchronize()
    assert util.maxdiff(o1, _nhwc(ref_relu)) <= 2e-5
    assert util.maxdiff(o2, _nhwc(ref_skip)) <= 2e-5
    assert util.maxdiff(s2.float(), o2) <= 1e-5


def test_maxpool_corr_im2col():
    dev = _dev()
    from nopesac_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, 30, 40, generator=g)
    p = ops.maxpool2_planes(_nhwc(x).to(dev), 2, 30, 40)
    assert util.maxdiff(p.float(), _nhwc(F.max_pool2d(x, 2, 2))) <= 1e-6
    B, C, H, W = 2, 256, 15, 20
    f1, f2 = torch.randn(B, C, H, W, generator=g) * 0.2, torch.randn(B, C, H, W, generator=g) * 0.2
    aff = restate.compute_corr_softmax(f1.double(), f2.double())            # [B, HW, H, W]
    got = ops.corr_softmax(_nhwc(f1).to(dev), _nhwc(f2).to(dev), B, H, W)
    torch.cuda.synchronize()
    assert got.hi.shape == (B * H * W, 320)
    assert util.maxdiff(got.float()[:, :H * W], _nhwc(aff)) <= 1e-6
    assert float(got.float()[:, H * W:].abs().max()) == 0.0
    for stride in (1, 2):
        y = torch.randn(2, 128, 15, 20, generator=g)
        cols, Ho, Wo = ops.im2col3x3_planes(_nhwc(y).to(dev), 2, 15, 20, stride)
        ref = F.unfold(y, 3, padding=1, stride=stride).reshape(2, 128, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(2 * Ho * Wo, 9 * 128)
        assert (Ho, Wo) == ((15 - 1) // stride + 1, (20 - 1) // stride + 1)
        assert util.maxdiff(cols.float(), ref) <= 1e-6

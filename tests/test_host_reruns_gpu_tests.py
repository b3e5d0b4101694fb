"""CPU: bodies of selected GPU tests re-run on the HOST — `_dev()` / `_gpu()` return the CPU device, `torch.cuda.synchronize`
is a no-op and the library handle is the host build (tests/host_fixture.py: plain-SIMT kernel sources + functional stand-ins for
the tensor-engine entry points).  Same assertions, same tolerances as on the device; what is exercised here is the kernel source
of the SIMT kernels and all Python glue, not the tcgen05 kernels."""
import pytest
import torch

from tests import host_fixture, test_gpu_eval, test_gpu_parity, test_gpu_pixel


@pytest.fixture()
def on_host(monkeypatch):
    host_fixture.install(monkeypatch, with_tensor_standins=True)
    cpu = torch.device("cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    for mod, name in ((test_gpu_pixel, "_dev"), (test_gpu_parity, "_gpu"), (test_gpu_eval, "_dev")):
        monkeypatch.setattr(mod, name, lambda: cpu)
    return cpu


@pytest.mark.parametrize("N,C,H,W,Cout", [(3, 64, 15, 20, 128), (2, 128, 7, 9, 64)])
def test_pixel_planes_and_conv_glue(on_host, N, C, H, W, Cout):
    test_gpu_pixel.test_conv3x3_implicit_gemm(N, C, H, W, Cout)          # nchw_to_planes from source; conv = stand-in


def test_pixel_groupnorm_upsample_add(on_host):
    test_gpu_pixel.test_groupnorm_upsample_add()


def test_pixel_maxpool_corr_im2col(on_host):
    test_gpu_pixel.test_maxpool_corr_im2col()


def test_parity_geo_sequence_kernel(on_host):
    test_gpu_parity.test_geo_sequence_kernel_on_oracle_inputs()


def test_parity_linear_kernel(on_host):
    test_gpu_parity.test_linear_kernel_against_torch()

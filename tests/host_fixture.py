"""Shared pytest fixture for the host execution of the product (CPU suite): swaps the library handle of nopesac_b200._lib for
the host build of the plain-SIMT kernel sources (+ functional stand-ins for the tensor-engine entry points when asked), and the
two CUDA-isms of nopesac_b200.ops (the CUDA-tensor check and the stream getter).  Test-side only — the product keeps failing
loudly on CPU tensors (tests/test_cabi.py, tests/test_oracle_planes.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_host"))

import run as simt_run  # noqa: E402

from nopesac_b200 import _lib, ops  # noqa: E402


def install(monkeypatch, with_tensor_standins: bool):
    L = simt_run.build(simt_run.SIMT_SOURCES + (simt_run.HOST_ONLY_SOURCES if with_tensor_standins else []),
                       extra_cpp=("tc_standin.cpp",) if with_tensor_standins else ())
    for name, (res, args) in _lib._SIGNATURES.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_lib, "_lib", L)

    def chk(t, name, dtype=torch.float32):
        if t.dtype != dtype:
            raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
        return t

    monkeypatch.setattr(ops, "_chk", chk)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    return ops

"""TEST INFRASTRUCTURE — not product code.

Loads the *unmodified* reference hot-path modules straight from /root/reference
(read-only; nothing is copied) so that

  * oracle/restate.py can be validated against the real reference code, and
  * tests/golden/make_golden.py can generate the committed golden fixtures.

The reference cannot travel to the GPU box (/root/reference does not exist
there), so nothing that runs under `-m gpu`, smoke() or bench.py imports this
module.  `available()` says whether the reference tree is present.

What the harness has to fake:
  detectron2.utils.registry.Registry, detectron2.config.configurable,
  detectron2.layers.{Conv2d, ShapeSpec, get_norm, FrozenBatchNorm2d},
  fvcore.nn.weight_init.{c2_xavier_fill, c2_msra_fill}, an empty `quaternion`
  module; and on CPU `torch.Tensor.cuda` must be the identity because the
  reference hard-codes `.cuda()` in matching_head.py:274-301.
"""
from __future__ import annotations

import contextlib
import functools
import importlib.util
import inspect
import io
import os
import sys
import types

import torch
from torch import nn
from torch.nn import functional as F

REF_ROOT = os.environ.get("NSAC_REFERENCE_ROOT", "/root/reference")
_MODELING = os.path.join(REF_ROOT, "NopeSAC_Net", "modeling")
_PKG = "_nsac_refpkg"


def available() -> bool:
    return os.path.isfile(os.path.join(_MODELING, "camera_net", "camera_head.py"))


# --------------------------------------------------------------------------- stubs
class _Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._obj_map[name]


def _configurable(init_func=None, *, from_config=None):
    assert init_func is not None and from_config is None

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        fc = type(self).from_config
        first = args[0] if args else kwargs.get("cfg")
        if first is not None and hasattr(first, "MODEL"):
            explicit = fc(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)
    return wrapped


class _ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class _Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv -> optional norm -> optional activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    if norm == "GN":
        return nn.GroupNorm(32, out_channels)
    if norm == "BN":
        return nn.BatchNorm2d(out_channels)
    raise NotImplementedError(norm)


def _c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def _c2_msra_fill(module):
    nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__nsac_stub__ = True
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    try:  # use the real thing if it is ever installed
        import detectron2  # noqa: F401
        import fvcore  # noqa: F401
    except Exception:
        mod("detectron2")
        mod("detectron2.utils")
        mod("detectron2.utils.registry", Registry=_Registry)
        mod("detectron2.config", configurable=_configurable)
        mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=_ShapeSpec, get_norm=_get_norm,
            FrozenBatchNorm2d=nn.BatchNorm2d)
        mod("fvcore")
        mod("fvcore.nn")
        mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill, c2_msra_fill=_c2_msra_fill)
        sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]
    if "quaternion" not in sys.modules:
        try:
            import quaternion  # noqa: F401
        except Exception:
            mod("quaternion")


class _Ref:
    pass


_CACHE = None


def load():
    """Returns a namespace with the reference modules: camera_modules, camera_head, gnn, matching_head."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    pkg(_PKG, _MODELING)
    pkg(_PKG + ".camera_net", os.path.join(_MODELING, "camera_net"))
    pkg(_PKG + ".matching_net", os.path.join(_MODELING, "matching_net"))
    pkg(_PKG + ".transformer", os.path.join(_MODELING, "transformer"))

    def load_file(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(_MODELING, relpath))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    ref = _Ref()
    ref.camera_modules = load_file(_PKG + ".camera_net.camera_modules", "camera_net/camera_modules.py")
    ref.camera_head = load_file(_PKG + ".camera_net.camera_head", "camera_net/camera_head.py")
    ref.gnn = load_file(_PKG + ".transformer.gnn", "transformer/gnn.py")
    ref.matching_head = load_file(_PKG + ".matching_net.matching_head", "matching_net/matching_head.py")
    ref.ShapeSpec = sys.modules["detectron2.layers"].ShapeSpec
    _CACHE = ref
    return ref


# --------------------------------------------------------------------------- cfg
class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(num_queries=50, out_cam_type="soft", match_threshold=0.2):
    """Attribute-dict with exactly the keys the hot path reads (SURVEY.md §5 'Config / flags'),
    set to configs/inference_mp3d.yaml + config/config.py defaults."""
    A = AttrDict
    return A(
        MODEL=A(
            EMBEDDING_ON=True, MASK_ON=True, DEVICE="cpu",
            CAMERA_HEAD=A(
                NAME="PlaneCameraHead", RAND_ON=False, CAM_REC_ON=True, REFINE_ON=True,
                INFERENCE_SP_TOPCAM_ON=False, INFERENCE_SP_TOPCAM_PATH="",
                PLANE_CAM_WEIGHT=1.0, PLANE_CAM_WEIGHT_PREDPLANE=0.1, INITIAL_CAM_WEIGHT=1.0,
                INFERENCE_OUT_CAM_TYPE=out_cam_type, WARP_PLANE_IN_CAM_REF_ON=True,
            ),
            SEM_SEG_HEAD=A(
                NUM_OBJECT_QUERIES=num_queries, IN_FEATURES=["res2", "res3", "res4", "res5"],
                CONVS_DIM=128, MASK_DIM=256, NORM="GN",
            ),
            MATCHING_HEAD=A(NAME="", OFFSET_MULTIPLIER=4.0, NORMAL_MULTIPLIER=8.0),
        ),
        TEST=A(
            MATCHING_SCORE_THRESHOLD=match_threshold,
            POSE_REFINEMENT_WITH_GT_MATCHERS=False,
            POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS=False,
            POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS_OFFSET_SCALE=0.1,
            POSE_REFINEMENT_WITH_GT_NOISE_MATCHERS_NORMAL_SCALE=10.0,
        ),
    )


def input_shape():
    S = load().ShapeSpec
    return {"res2": S(channels=256, stride=4), "res3": S(channels=512, stride=8),
            "res4": S(channels=1024, stride=16), "res5": S(channels=2048, stride=32)}


@contextlib.contextmanager
def cpu_patch():
    """`.cuda()` -> identity while the reference runs on CPU (matching_head.py:274-301),
    and swallow the head's per-call print() (camera_head.py:510)."""
    had = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            yield
    finally:
        torch.Tensor.cuda = had


def build_heads(num_queries=50, out_cam_type="soft", seed=40):
    """Reference PlaneCameraHead + MatchingHead, eval mode, reference initialisers under
    torch.manual_seed(40) (siamese_planeTR.py:51)."""
    ref = load()
    cfg = make_cfg(num_queries=num_queries, out_cam_type=out_cam_type)
    torch.manual_seed(seed)
    head = ref.camera_head.PlaneCameraHead(cfg, input_shape()).eval()
    match = ref.matching_head.MatchingHead(cfg).eval()
    return head, match, cfg


def private(head, name):
    """Name-mangled private stage of the reference head, e.g. private(head, 'inference_PlaneCamRefHead')."""
    return getattr(head, "_PlaneCameraHead__" + name)

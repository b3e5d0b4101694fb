"""TEST INFRASTRUCTURE — loads the reference's UNMODIFIED `PlaneTRHead` (planeTR_net/planeTR_head.py with transformer/transformer.py
and position_encoding.py) straight from /root/reference under the detectron2 / fvcore stubs of oracle/ref_loader.py, to pin
oracle/planeTR_restate.py and to generate tests/golden/planetr_*.pt.  Build container only (see ref_loader)."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from oracle import ref_loader

_CACHE = None


def available() -> bool:
    return os.path.isfile(os.path.join(ref_loader._MODELING, "planeTR_net", "planeTR_head.py"))


def load():
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    ref_loader.load()                       # installs the stubs and the synthetic package (incl. `.transformer`)
    pkg = ref_loader._PKG
    m = types.ModuleType(pkg + ".planeTR_net")
    m.__path__ = [os.path.join(ref_loader._MODELING, "planeTR_net")]
    sys.modules[pkg + ".planeTR_net"] = m

    def load_file(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(ref_loader._MODELING, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    load_file(pkg + ".transformer.position_encoding", "transformer/position_encoding.py")
    load_file(pkg + ".transformer.transformer", "transformer/transformer.py")
    _CACHE = load_file(pkg + ".planeTR_net.planeTR_head", "planeTR_net/planeTR_head.py")
    return _CACHE


def make_cfg(num_queries=50, enc_layers=6, dec_layers=6):
    A = ref_loader.AttrDict
    return A(MODEL=A(DEPTH_ON=False,
                     SEM_SEG_HEAD=A(NAME="PlaneTRHead", NUM_CLASSES=1, PARAM_ON=True, CENTER_ON=True, DEEP_SUPERVISION=True,
                                    MASK_DIM=256, HIDDEN_DIM=256, NUM_OBJECT_QUERIES=num_queries, NHEADS=8,
                                    ENC_LAYERS=enc_layers, DEC_LAYERS=dec_layers)))


def build_head(num_queries=50, enc_layers=6, dec_layers=6):
    mod = load()
    S = sys.modules["detectron2.layers"].ShapeSpec
    shape = {"res2": S(channels=256, stride=4), "res3": S(channels=512, stride=8), "res4": S(channels=1024, stride=16),
             "res5": S(channels=2048, stride=32)}
    return mod.PlaneTRHead(make_cfg(num_queries, enc_layers, dec_layers), shape).eval()

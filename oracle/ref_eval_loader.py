"""TEST INFRASTRUCTURE — not product code.

Executes the reference's own camera-evaluation code without importing its module (evaluation/mp3d_evaluation.py pulls in
detectron2, pycocotools, sklearn pickles ... at import time): the two definitions are cut out of the *unmodified* source file
with `ast` and compiled as they are.  Nothing is copied into the repo.  `available()` says whether the reference is present
(it is not on the GPU box).
"""
from __future__ import annotations

import ast
import os
import types

import numpy as np

REF_ROOT = os.environ.get("NSAC_REFERENCE_ROOT", "/root/reference")
_FILE = os.path.join(REF_ROOT, "NopeSAC_Net", "evaluation", "mp3d_evaluation.py")


def available() -> bool:
    return os.path.isfile(_FILE)


def _extract(names):
    with open(_FILE) as f:
        src = f.read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            found[node.name] = ast.get_source_segment(src, node)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"reference evaluation code not found: {sorted(missing)}")
    return found


def load():
    """Returns (angle_error_vec, eval_camera_reg(predictions, camera_name) -> metrics dict) backed by the reference source."""
    import textwrap
    srcs = _extract(["angle_error_vec", "_eval_camera_reg"])
    ns = {"np": np, "create_small_table": lambda d: str(d)}
    exec(compile(srcs["angle_error_vec"], _FILE, "exec"), ns)
    exec(compile(textwrap.dedent(srcs["_eval_camera_reg"]), _FILE, "exec"), ns)

    class _Log:
        def info(self, *a, **k):
            pass

    def eval_camera_reg(predictions, camera_name="camera"):
        cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(CAMERA_HEAD=types.SimpleNamespace(INFERENCE_OUT_CAM_TYPE="soft")))
        self = types.SimpleNamespace(_logger2=_Log(), cfg=cfg, _results={})
        ns["_eval_camera_reg"](self, predictions, camera_name)
        return self._results

    return ns["angle_error_vec"], eval_camera_reg


def load_get_optimized_dict():
    """The reference's `MP3DEvaluator.get_optimized_dict` (mp3d_evaluation.py:259-313) as a plain function of the predictions."""
    import textwrap
    src = _extract(["get_optimized_dict"])["get_optimized_dict"]
    ns = {"np": np}
    exec(compile(textwrap.dedent(src), _FILE, "exec"), ns)
    return lambda predictions: ns["get_optimized_dict"](types.SimpleNamespace(), predictions)

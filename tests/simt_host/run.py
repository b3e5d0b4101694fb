"""TEST INFRASTRUCTURE — builds a plain-SIMT .cu file of nopesac_b200/csrc for the HOST (tests/simt_host/cuda_runtime.h) and
returns it as a ctypes library, so the kernel source can be executed against the oracle in a container without a GPU."""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "nopesac_b200", "csrc")


def _split_top_level(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src: str) -> str:
    """`kernel<<<grid, block, smem, stream>>>(args);` -> `SIMT_LAUNCH(grid, block, kernel(args));`"""
    pat = re.compile(r"(\w+(?:<[\w\s,]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)

    def sub(m):
        cfg = _split_top_level(m.group(2))
        return f"SIMT_LAUNCH({cfg[0]}, {cfg[1]}, {m.group(1)}({m.group(3)}));"

    out, n = pat.subn(sub, src)
    if n == 0:
        raise RuntimeError("no kernel launch found")
    return out


def build(cu_name: str) -> ctypes.CDLL:
    with open(os.path.join(CSRC, cu_name)) as f:
        src = rewrite_launches(f.read())
    with open(os.path.join(HERE, "cuda_runtime.h")) as f, open(os.path.join(HERE, "simt_runtime.cpp")) as g:
        key = hashlib.sha1((src + f.read() + g.read()).encode()).hexdigest()[:16]
    out_dir = os.path.join(tempfile.gettempdir(), "nsac_simt_host")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"{os.path.splitext(cu_name)[0]}_{key}.so")
    if not os.path.exists(lib):
        gen = os.path.join(out_dir, f"{os.path.splitext(cu_name)[0]}_{key}.cpp")
        with open(gen, "w") as f:
            f.write(src)
        # the generated file sits outside csrc/: -I csrc for "common.cuh", which includes "../../include/nopesac_b200.h"
        # relative to ITS OWN directory, so the real header is used.
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w", "-x", "c++",
               "-I", HERE, "-I", CSRC, gen, os.path.join(HERE, "simt_runtime.cpp"), "-o", lib + ".tmp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("simt_host build failed:\n" + res.stderr[-4000:])
        os.replace(lib + ".tmp", lib)
    return ctypes.CDLL(lib)

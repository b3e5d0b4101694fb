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


def rewrite_launches(src: str, allow_none: bool = False) -> str:
    """`kernel<<<grid, block, smem, stream>>>(args);` -> `SIMT_LAUNCH(grid, block, smem, kernel(args));`"""
    pat = re.compile(r"(\w+(?:<[\w\s,]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)

    def sub(m):
        cfg = _split_top_level(m.group(2))
        smem = cfg[2] if len(cfg) > 2 else "0"
        return f"SIMT_LAUNCH({cfg[0]}, {cfg[1]}, {smem}, {m.group(1)}({m.group(3)}));"

    out, n = pat.subn(sub, src)
    if n == 0 and not allow_none:
        raise RuntimeError("no kernel launch found")
    # `extern __shared__ [__align__(16)] T name[];` -> pointer to the running block's dynamic shared memory
    out = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                 lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(simt::dyn_smem());", out)
    return out


def build(cu_names, extra_cpp=()) -> ctypes.CDLL:
    """One host library from one or several csrc/*.cu files (a str or a list of names) plus test-side .cpp files of this
    directory (`extra_cpp`, e.g. the tensor-engine stand-ins)."""
    if isinstance(cu_names, str):
        cu_names = [cu_names]
    srcs = {}
    for name in cu_names:
        with open(os.path.join(CSRC, name)) as f:
            srcs[name] = rewrite_launches(f.read(), allow_none=name in HOST_ONLY_SOURCES)
    h = hashlib.sha1()
    for name in sorted(srcs):
        h.update(srcs[name].encode())
    for dep in ("cuda_runtime.h", "cuda_fp16.h", "cuda_bf16.h", "simt_runtime.cpp") + tuple(extra_cpp):
        with open(os.path.join(HERE, dep), "rb") as f:
            h.update(f.read())
    with open(os.path.join(CSRC, "common.cuh"), "rb") as f:
        h.update(f.read())
    # NSAC_SIMT_SANITIZE=address|thread: instrumented build; run python under LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    # (or libtsan.so) — scripts/simt_sanitize.sh
    san = os.environ.get("NSAC_SIMT_SANITIZE", "")
    h.update(san.encode())
    key = h.hexdigest()[:16]
    out_dir = os.path.join(tempfile.gettempdir(), "nsac_simt_host")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"simt_{key}.so")
    if not os.path.exists(lib):
        gens = []
        for name, src in srcs.items():
            gen = os.path.join(out_dir, f"{os.path.splitext(name)[0]}_{key}.cpp")
            with open(gen, "w") as f:
                f.write(src)
            gens.append(gen)
        # the generated files sit outside csrc/: -I csrc for "common.cuh", which includes "../../include/nopesac_b200.h"
        # relative to ITS OWN directory, so the real header is used.
        opt = ["-O1", "-g", f"-fsanitize={san}", "-fno-omit-frame-pointer"] if san else ["-O2"]
        cmd = ["g++"] + opt + ["-std=c++17", "-fPIC", "-shared", "-pthread", "-fopenmp", "-ffp-contract=off", "-w",
               "-I", HERE, "-I", CSRC] + gens + [os.path.join(HERE, c) for c in ("simt_runtime.cpp",) + tuple(extra_cpp)] + \
              ["-o", lib + ".tmp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("simt_host build failed:\n" + res.stderr[-4000:])
        os.replace(lib + ".tmp", lib)
    return ctypes.CDLL(lib)


# host-side orchestration only (no kernels): needs the tensor-engine entry points, i.e. the stand-ins of tc_standin.cpp
HOST_ONLY_SOURCES = ["forward.cu"]
SIMT_SOURCES = ["dense.cu", "geo.cu", "matcher.cu", "score.cu", "evaluate.cu", "planes.cu", "pixel.cu", "backbone.cu", "planetr.cu"]   # no TMA / tcgen05 / inline PTX

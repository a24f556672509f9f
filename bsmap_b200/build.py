"""Build libbsmap_b200.so (CUDA kernels + C ABI + host text layer) in-tree for sm_100a.

    python -m bsmap_b200.build          # incremental: rebuilds only when a source is newer
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbsmap_b200.so")
CLI = os.path.join(HERE, "bsmap")
METH_CLI = os.path.join(HERE, "methratio")
SOURCES = ["bsx_index.cu", "bsx_map_se.cu", "bsx_map_se_wide.cu", "bsx_map_se_rrbs.cu", "bsx_map_pe.cu", "bsx_map_pe_rrbs.cu", "bsx_map_pe_wide.cu", "bsx_api.cu", "bsx_meth.cu", "bsx_format.cpp", "bsx_reads.cpp", "bsx_bam.cpp", "bsx_cli.cpp", "bsx_methratio_cli.cpp"]
HEADERS = ["bsx_common.cuh", "bsx_prep.cuh", "bsx_internal.h", "bsx_map.cuh", "bsx_map_impl.cuh", os.path.join("..", "..", "include", "bsmap_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI) or not os.path.exists(METH_CLI):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(CLI))
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, s))]
    deps.append(os.path.join(CSRC, "bsmap_main.cpp"))
    deps.append(os.path.join(CSRC, "methratio_main.cpp"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(
                [os.path.getmtime(s)] + [os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS]):
            continue
        cmd = [nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {s}")
    link = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread", "-lz"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    for exe, src in ((CLI, "bsmap_main.cpp"), (METH_CLI, "methratio_main.cpp")):
        main_cpp = os.path.join(CSRC, src)
        if os.path.exists(main_cpp):
            r = subprocess.run(["g++", "-O2", "-o", exe, main_cpp, "-L" + HERE, "-lbsmap_b200", "-Wl,-rpath,$ORIGIN"],
                               capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"{os.path.basename(exe)} CLI link failed")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

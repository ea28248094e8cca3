"""In-tree build recipes (nvcc for sm_100a, g++ for the host-only helpers).

Artefacts land next to this file so that they travel to the GPU box with the
repo snapshot:
  libspaghetti_gpu.so  the product: CUDA kernels + the C ABI of include/spaghetti.h
  libss_synth.so       synthetic workload generators
  libspaghetti_host.so C++ mirror of the Go API over table snapshots, calls the C ABI
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
GPU_LIB = PKG / "libspaghetti_gpu.so"
SYNTH_LIB = PKG / "libss_synth.so"
HOST_LIB = PKG / "libspaghetti_host.so"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _host_cxx() -> str:
    # $CXX in this image points at a g++ build without libgomp.spec; the one on
    # PATH is complete.
    for cand in ("/usr/bin/g++", shutil.which("g++") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no g++ found")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc") or "", "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(str(c) for c in cmd), file=sys.stderr)
    subprocess.run([str(c) for c in cmd], check=True)


def build_synth(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / "synth" / "synth.cpp", CSRC / "synth" / "synth.h"]
    if force or _stale(SYNTH_LIB, srcs):
        _run([_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-Wall", "-shared",
              "-o", SYNTH_LIB, srcs[0]], verbose)
    return SYNTH_LIB


def gpu_sources():
    cu = sorted(CSRC.glob("*.cu"))
    hdr = (sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.inc")) +
           [ROOT / "include" / "spaghetti.h"])
    return cu, hdr


def build_gpu(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a (one object per file, in parallel) into one shared library."""
    from concurrent.futures import ThreadPoolExecutor
    cu, hdr = gpu_sources()
    if not (force or _stale(GPU_LIB, cu + hdr)):
        return GPU_LIB
    obj_dir = PKG / "_obj"
    obj_dir.mkdir(exist_ok=True)
    base = [_nvcc(), *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-fvisibility=hidden",
            "-I", ROOT / "include", "-I", CSRC, "-ccbin", _host_cxx()]
    if ptxas_info:
        base += ["-Xptxas", "-v"]
    objs, jobs = [], []
    for src in cu:
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdr):
            jobs.append(base + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as pool:
        list(pool.map(lambda c: _run(c, verbose), jobs))
    _run([_nvcc(), *NVCC_ARCH, "-shared", "-ccbin", _host_cxx(), "-o", GPU_LIB, *objs, "-ldl"], verbose)
    return GPU_LIB


def build_host(force: bool = False, verbose: bool = False) -> Path:
    """C++ mirror of the reference's Go API over table snapshots (csrc/host), on top of the C ABI."""
    srcs = [CSRC / "host" / "host_mirror.cpp", CSRC / "host" / "host_mirror.h", ROOT / "include" / "spaghetti.h"]
    build_gpu(force=False, verbose=verbose)
    if force or _stale(HOST_LIB, srcs + [GPU_LIB]):
        _run([_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-Wall", "-fvisibility=hidden", "-shared", "-I", ROOT / "include",
              "-o", HOST_LIB, srcs[0], "-L", PKG, "-lspaghetti_gpu", "-Wl,-rpath,$ORIGIN"], verbose)
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_synth(force, verbose), build_gpu(force, verbose), build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)

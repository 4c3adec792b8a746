"""Builds libqarbom_b200.so (sm_100a only) in-tree with nvcc; nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libqarbom_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    out = [os.path.join(HERE, "..", "include", "qarbom_b200.h")]
    for f in sorted(os.listdir(SRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(SRC, f))
    return out


def _nccl_device_api_flags():
    """nvls_setup.cu is compiled against NCCL's own headers when the device API (NCCL >= 2.28: nccl_device/*.h, shipped in
    the nvidia-nccl wheel next to the library torch loads) is present; without them the NVSwitch-multicast exchange reports
    itself unavailable at run time and the NCCL collectives are used."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    except (ImportError, ValueError):
        roots = []
    for r in roots:
        inc = os.path.join(r, "include")
        if os.path.isfile(os.path.join(inc, "nccl_device", "impl", "comm__types.h")) and os.path.isfile(os.path.join(inc, "nccl.h")):
            return ["-DQB_HAVE_NCCL_DEVICE_API", "-I" + inc, "-diag-suppress", "821"]
    return []


def nccl_runtime_library():
    """Path of the libnccl.so.2 that matches those headers (None when the wheel is absent)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    except (ImportError, ValueError):
        roots = []
    for r in roots:
        p = os.path.join(r, "lib", "libnccl.so.2")
        if os.path.isfile(p):
            return p
    return None


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *_nccl_device_api_flags(), os.path.join(SRC, "qb_api.cu"), os.path.join(SRC, "nvls_setup.cu"),
           "-o", LIB, "-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

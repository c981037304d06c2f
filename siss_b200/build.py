"""Build libsiss_b200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a C ABI).

Run as ``python -m siss_b200.build`` or through ``__graft_entry__.build()``. The shared object is
written next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libsiss_b200.so"
HASH_PATH = PKG_DIR / "libsiss_b200.so.srchash"
SOURCES = ["rowwise.cu", "wmse.cu", "combine.cu", "p2p.cu", "nvls.cu", "ce.cu", "stats.cu", "optim.cu", "multitensor.cu", "membership.cu", "rng.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--shared",
    "--threads", "0",          # compile the translation units in parallel (one per core)
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsiss_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def source_hash() -> str:
    """SHA-256 over everything the library is compiled from (sources, headers, flags). Stored next to the built
    library (``libsiss_b200.so.srchash``) so that a stale ``.so`` is detected by CONTENT — file times do not survive
    the snapshot copy to a GPU box, and an edit that leaves the ABI version unchanged must still trigger a rebuild."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted([CSRC / s for s in SOURCES if (CSRC / s).exists()] + list(CSRC.glob("*.cuh")))
    deps.append(PKG_DIR.parent / "include" / "siss_b200.h")
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not LIB_PATH.exists() or not HASH_PATH.exists():
        return True
    try:
        return HASH_PATH.read_text().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    # build next to the target and rename: atomic, so concurrent builders (one per rank under torchrun)
    # can only ever duplicate work, never expose a half-written library
    tmp = LIB_PATH.with_name(f"{LIB_PATH.name}.tmp.{os.getpid()}")
    cmd += ["-o", str(tmp), *srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        try:
            tmp.unlink()
        except OSError:
            pass
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    if verbose:
        sys.stderr.write(proc.stderr)
    os.replace(tmp, LIB_PATH)
    htmp = HASH_PATH.with_name(f"{HASH_PATH.name}.tmp.{os.getpid()}")
    htmp.write_text(source_hash() + "\n")
    os.replace(htmp, HASH_PATH)
    return LIB_PATH


# --------------------------------------------------------------------------------------------------
# thin torch extension over the C ABI (csrc/torch_ext.cpp): torch.ops.siss_b200.*
# --------------------------------------------------------------------------------------------------
EXT_SRC = CSRC / "torch_ext.cpp"
EXT_PATH = PKG_DIR / "_siss_torch_ext.so"
EXT_HASH_PATH = PKG_DIR / "_siss_torch_ext.so.srchash"


def ext_source_hash() -> str:
    import hashlib
    import torch
    h = hashlib.sha256()
    h.update(EXT_SRC.read_bytes())
    h.update((PKG_DIR.parent / "include" / "siss_b200.h").read_bytes())
    h.update(torch.__version__.encode())
    return h.hexdigest()


def ext_needs_build() -> bool:
    if not EXT_PATH.exists() or not EXT_HASH_PATH.exists():
        return True
    try:
        return EXT_HASH_PATH.read_text().strip() != ext_source_hash()
    except OSError:
        return True


def build_torch_ext(force: bool = False) -> Path:
    """g++ (no nvcc: the file contains no device code) against the torch headers; links libsiss_b200.so through
    ``$ORIGIN`` so the pair can be moved together. Atomic like :func:`build`."""
    build()                                            # the C-ABI library it forwards to
    if not force and not ext_needs_build():
        return EXT_PATH
    import torch
    from torch.utils import cpp_extension as ce
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    inc = ce.include_paths("cuda") if hasattr(ce, "include_paths") else []
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    inc = list(inc) + [os.path.join(cuda_home, "include")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    tmp = EXT_PATH.with_name(f"{EXT_PATH.name}.tmp.{os.getpid()}")
    cmd = [cxx, "-shared", "-fPIC", "-O2", "-std=c++17", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", *[f"-I{i}" for i in inc], str(EXT_SRC), "-o", str(tmp),
           f"-L{PKG_DIR}", "-l:libsiss_b200.so", f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        try:
            tmp.unlink()
        except OSError:
            pass
        raise RuntimeError(f"{cxx} failed ({proc.returncode}) building the torch extension")
    os.replace(tmp, EXT_PATH)
    htmp = EXT_HASH_PATH.with_name(f"{EXT_HASH_PATH.name}.tmp.{os.getpid()}")
    htmp.write_text(ext_source_hash() + "\n")
    os.replace(htmp, EXT_HASH_PATH)
    return EXT_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    if "--no-ext" not in sys.argv:
        print(build_torch_ext(force="--force" in sys.argv))

"""CPU: the C-ABI library loads and exports every symbol include/siss_b200.h declares, the ctypes
table matches the header arity, and the product path refuses to run without a CUDA device (no
compute calls are made here)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "siss_b200.h").read_text()


def declared_functions():
    src = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"^(?:int|int64_t|const char\*)\s+(siss_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.M | re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_header_parses():
    fns = declared_functions()
    assert len(fns) >= 17
    for must in ("siss_add_noise_pair", "siss_add_noise_mixture", "siss_mixture_weights", "siss_wmse_fwd_bwd",
                 "siss_dual_mse_fwd_bwd", "siss_norm3", "siss_combine"):
        assert must in fns


def test_library_builds_and_exports_every_declared_symbol():
    from siss_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/siss_b200.h but not exported"


def test_ctypes_table_matches_header():
    from siss_b200 import _lib
    fns = declared_functions()
    assert set(_lib.SIGNATURES) == set(fns)
    for name, nargs in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
    lib = _lib.load()
    assert lib.siss_abi_version() == _lib.ABI_VERSION
    assert lib.siss_row_workspace_bytes(64) > 0 and lib.siss_norm3_workspace_bytes() > 0
    assert "not sm_100" in _lib.error_string(-3)


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly, not silently computed some other way."""
    from siss_b200 import ops
    from siss_b200._lib import SissLibraryError
    from siss_b200.losses import DDPMDeletionLoss
    from siss_b200.scheduler import SissDDPMScheduler
    x = torch.zeros(2, 1, 4, 4)
    t = torch.zeros(2, dtype=torch.long)
    sched = SissDDPMScheduler()
    with pytest.raises(SissLibraryError):
        sched.add_noise(x, x, t)
    with pytest.raises(SissLibraryError):
        ops.norm3(torch.zeros(8), torch.zeros(8))
    gamma, sigma = sched.gamma_sigma("cpu")
    loss = DDPMDeletionLoss(gamma, sigma)
    d = {"og_latents": x, "noisy_latents": x}
    with pytest.raises(SissLibraryError):
        loss.importance_sampling_with_mixture(lambda *a, **k: (x,), t, x, {}, d, d, lambd=0.5)
    with pytest.raises(SissLibraryError):
        loss.naive_del(lambda *a, **k: (x,), t, x, {}, d, d)


def test_product_does_not_import_oracle():
    for p in (ROOT / "siss_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "siss_oracle" not in txt, p


def test_scheduler_tables_match_oracle():
    from oracle import siss_oracle as O
    from siss_b200.scheduler import SissDDPMScheduler
    s = SissDDPMScheduler()
    assert torch.equal(s.alphas_cumprod, O.make_alphas_cumprod())
    s = SissDDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    assert torch.equal(s.alphas_cumprod, O.make_alphas_cumprod(beta_start=0.00085, beta_end=0.012,
                                                                beta_schedule="scaled_linear"))
    g, sg = s.gamma_sigma("cpu")
    og, os_ = O.gamma_sigma(s.alphas_cumprod)
    assert torch.equal(g, og) and torch.equal(sg, os_)


def test_argument_validation_without_a_gpu():
    """Every entry point validates pointers / sizes / enums before it touches CUDA, so the error
    behaviour of the ABI can be checked on a CPU box: SISS_EINVAL (-1), never a crash or a launch."""
    from siss_b200 import _lib
    lib = _lib.load()
    N = None  # NULL
    one = ctypes.c_void_p(16)  # a non-null, 16-byte aligned dummy address: never dereferenced on these paths
    assert lib.siss_add_noise(N, one, one, one, 1000, one, 4, 16, 0, N) == -1
    assert lib.siss_add_noise(one, one, one, one, 0, one, 4, 16, 0, N) == -1            # T < 1
    assert lib.siss_add_noise(one, one, one, one, 1000, one, -1, 16, 0, N) == -1        # negative batch
    assert lib.siss_add_noise(one, one, one, one, 1000, one, 0, 16, 0, N) == 0          # empty batch: no-op
    assert lib.siss_add_noise(one, one, one, one, 1000, one, 4, 16, 7, N) == -2         # unknown dtype
    assert lib.siss_add_noise_pair(one, N, one, one, one, 1000, one, one, 4, 16, 0, N) == -1
    assert lib.siss_mixture_weights(one, one, one, one, one, one, one, one, 1000, 0.5, one, one, one, one, one, N,
                                    4, 16, 0, N) == -1                                   # missing workspace
    assert lib.siss_add_noise_mixture(one, one, one, one, one, one, one, one, 1000, 0.5, one, one, one, one, one, one,
                                      4, 0, 0, N) == -1                                  # D < 1
    assert lib.siss_wmse_fwd_bwd(one, 0, one, one, one, 0, one, one, one, 1000, one, one, 1.0, 1.0, one, N, one, one,
                                 one, 4, 16, N) == -1
    assert lib.siss_wmse_fwd_bwd(one, 1, one, one, one, 0, one, one, one, 1000, one, one, 1.0, 1.0, one, one, one, one,
                                 one, 4, 16, N) == -2                                    # bf16 pred with fp32 latents
    assert lib.siss_sqerr_fwd(one, 1, one, 2, one, N, 0.0, 16, N) == -2                  # bf16 x fp16 not compiled in
    assert lib.siss_sqerr_bwd(one, 0, one, 0, N, 0, N, 0, 0.0, 0, one, 16, N) == -1      # no upstream gradient at all
    assert lib.siss_norm3(N, one, 16, one, one, N) == -1
    assert lib.siss_combine(one, one, one, 16, one, 3, 1.0, 1.0, 0, N, N) == -1          # bad mode
    assert lib.siss_combine(one, one, one, 16, N, 0, 1.0, 1.0, 0, N, N) == -1            # sums3 missing
    assert lib.siss_combine_adamw(one, one, 16, one, 0, 1.0, 1.0, 0, one, one, one, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, N,
                                  N, N, 0.0, 1, N, N, N) == -1                           # step < 1 and no device counter
    assert lib.siss_combine_adamw(one, one, 16, one, 0, 1.0, 1.0, 0, one, one, one, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, N,
                                  N, one, 1.5, 1, N, N, N) == -1                         # EMA decay outside [0, 1]
    assert lib.siss_counter_add(N, 1, N) == -1
    assert lib.siss_dual_mse_rng_fwd_bwd(one, one, 0, one, 0, 1, 1 << 62, N, 0, 1.0, 1.0, one, one, N, one, one, one, 1, 16,
                                         N) == -1                                            # draw >= 2^62
    assert lib.siss_randn(one, 16, 0, 1, 1 << 62, N, 0, N) == -1
    arr8 = (ctypes.c_void_p * 8)(*([one.value] * 8))
    assert lib.siss_p2p_adamw_allgather(one, one, one, arr8, 1, 0, 16, 0, 1.0, 1.0, 0, one, one, 1e-3, 0.9, 0.999, 1e-8,
                                        0.0, 1, N, N, N, 0.0, N, N) == -1                    # world < 2
    assert lib.siss_p2p_adamw_allgather(one, one, one, arr8, 2, 0, 16, 2, 1.0, 1.0, 0, one, one, 1e-3, 0.9, 0.999, 1e-8,
                                        0.0, 1, N, N, N, 0.0, N, N) == -1                    # single-term mode not offered
    assert lib.siss_membership_add_noise(one, one, one, one, 1000, 1000, one, one, 0, 1, 1, 16, 0, N) == -1   # t >= T
    assert lib.siss_membership_add_noise(one, one, one, one, 1000, 5, one, one, 0, 1, 0, 16, 0, N) == -1      # n_noise < 1
    assert lib.siss_membership_sqerr(one, one, one, 0, one, one, N, 0, 1, 1, 16, N) == -1                     # no workspace
    assert lib.siss_membership_sqerr(one, one, one, 7, one, one, one, 0, 1, 1, 16, N) == -2                   # dtype
    assert lib.siss_batch_stats(one, one, one, one, 0, 16, one, N) == -1                 # B < 1
    arr = (ctypes.c_void_p * 8)(*[16 * (i + 1) for i in range(8)])
    assert lib.siss_p2p_reduce_norm3(arr, arr, arr, 3, 0, 16, one, one, one, 0, one, N) == -2   # world must be 2, 4 or 8
    assert lib.siss_p2p_reduce_norm3(arr, arr, arr, 2, 2, 16, one, one, one, 0, one, N) == -1   # rank out of range
    assert lib.siss_p2p_combine_allgather(one, one, one, arr, 2, 0, 18, 0, 1.0, 1.0, 0, N, N) == -1  # shard % 4
    assert lib.siss_mt_norm3(one, one, one, one, 0, 1, one, one, N) == -1                # no tensors
    for code in (-1, -2, -3):
        assert _lib.error_string(code).startswith("siss:")


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/siss_b200.h must compile as strict C99 (and as C++), with no CUDA or
    torch headers, and a C program must link against libsiss_b200.so and call an entry point that needs no GPU."""
    import shutil
    import subprocess
    hdr = ROOT / "include" / "siss_b200.h"
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", str(hdr)], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", str(hdr)], check=True)
    txt = hdr.read_text()
    includes = [l.split()[1] for l in txt.splitlines() if l.startswith("#include")]
    assert set(includes) <= {"<stdint.h>", "<stddef.h>"}, includes            # no CUDA, no torch headers in the boundary
    from siss_b200 import _lib
    _lib.load()                                                   # builds the library if it is missing
    src = tmp_path / "main.c"
    src.write_text('#include <stdio.h>\n#include "siss_b200.h"\n'
                   'int main(void) {\n'
                   '  printf("%d %s|%lld\\n", siss_abi_version(), siss_error_string(SISS_EINVAL),\n'
                   '         (long long)siss_row_workspace_bytes(64));\n'
                   '  return siss_add_noise(0, 0, 0, 0, 1000, 0, 1, 1, SISS_F32, 0) == SISS_EINVAL ? 0 : 1;\n}\n')
    exe = tmp_path / "main"
    libdir = _lib.LIB_PATH.parent
    subprocess.run([gcc, "-std=c99", "-I", str(hdr.parent), str(src), "-o", str(exe), f"-L{libdir}", "-lsiss_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    ver, rest = out.split(" ", 1)
    assert int(ver) == _lib.ABI_VERSION and "invalid argument" in rest and int(rest.rsplit("|", 1)[1]) > 0

"""ctypes binding of libfringe_b200.so (the C ABI declared in include/fringe_b200.h).

The library is built in-tree by ``fringe_b200.build`` (``__graft_entry__.build()``).  If it is
missing the import fails loudly: there is no Python or CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRINGE_B200_LIB") or os.path.join(_HERE, "lib", "libfringe_b200.so")   # override: development builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fringe_b200.h")
PROF_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fringe_b200_prof.h")
PROF_LIB_PATH = os.path.join(_HERE, "lib", "libfringe_b200_prof.so")

OK, ERR_METHOD, ERR_ARGUMENT, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_MEMORY = range(7)
NMAP_KS2, NMAP_AD2 = 0, 1
EVD_EVD, EVD_MLE, EVD_STBAS = 0, 1, 2
VARIANT_EVD, VARIANT_PHASE_LINK = 0, 1


class FringeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"fringe_b200 status {status}: {message}")
        self.status = status


def declared_symbols(header: str | None = None) -> list[str]:
    """Every function name a header declares (include/fringe_b200.h by default; used by the ABI export test)."""
    text = open(header or HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fringe_[a-z0-9_]+)\s*\(", text)))


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fringe_b200.build` "
            "(fringe_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.fringe_abi_version.restype = i
    lib.fringe_device_count.argtypes = [C.POINTER(i)]
    lib.fringe_create.argtypes = [i, C.POINTER(vp)]
    lib.fringe_destroy.argtypes = [vp]
    lib.fringe_last_error.argtypes = [vp]
    lib.fringe_last_error.restype = C.c_char_p
    lib.fringe_status_string.argtypes = [i]
    lib.fringe_status_string.restype = C.c_char_p
    lib.fringe_synchronize.argtypes = [vp]
    lib.fringe_launch_count.argtypes = [vp]
    lib.fringe_launch_count.restype = C.c_int64
    lib.fringe_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.fringe_host_free.argtypes = [vp]
    lib.fringe_nulong.argtypes = [i, i]
    lib.fringe_ks2_critical_count.argtypes = [i, d, C.POINTER(i), C.POINTER(d)]
    lib.fringe_ad2_critical_sum.argtypes = [i, d, C.POINTER(d)]
    lib.fringe_ad2_sigma.argtypes = [i, C.POINTER(d)]
    lib.fringe_evd_max_bands.argtypes = [i, i]
    nmap_args = [vp, vp, vp, vp, i, i, i, i, i, i, d, vp, vp]
    lib.fringe_nmap_block.argtypes = nmap_args
    lib.fringe_nmap_block_device.argtypes = nmap_args + [vp]
    evd_args = [vp, vp, vp] + [i] * 12 + [vp, vp, vp]
    lib.fringe_evd_block.argtypes = evd_args
    lib.fringe_evd_block_device.argtypes = evd_args + [vp]
    lib.fringe_nmap_evd_block.argtypes = [vp, vp, vp, vp] + [i] * 6 + [d] + [i] * 7 + [vp] * 5
    lib.fringe_ampdispersion_block.argtypes = [vp, vp, vp, i, i, i, vp, vp]
    lib.fringe_ampdispersion_block_device.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp]
    lib.fringe_despeck_block.argtypes = [vp, vp, vp, vp] + [i] * 7 + [vp]
    lib.fringe_despeck_block_device.argtypes = [vp, vp, vp, vp] + [i] * 7 + [vp, vp]
    lib.fringe_sequential_halo.argtypes = [i, i, i]
    lib.fringe_sequential_block.argtypes = [vp, vp, vp] + [i] * 10 + [vp] * 6
    lib.fringe_calamp_block.argtypes = [vp, vp, vp, i, i, i, vp, vp]
    lib.fringe_integrate_ps.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, vp]
    lib.fringe_ps_coherence.argtypes = [vp, vp, vp, C.c_int64, C.c_float, vp]
    lib.fringe_cmul.argtypes = [vp, vp, vp, C.c_int64, vp]
    lib.fringe_cmul_device.argtypes = [vp, vp, vp, C.c_int64, vp, vp]
    # context-bound profiling hooks (include/fringe_b200_prof.h, group a)
    lib.fringe_evd_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.fringe_evd_phase_cycles.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.fringe_last_kernel_ms.argtypes = [vp, i, C.POINTER(C.c_float)]
    lib.fringe_prof_force_generic.argtypes = [vp, i]
    for name in declared_symbols():
        getattr(lib, name)          # AttributeError here = header / library mismatch
    return lib


lib = _load()


def prof_lib() -> C.CDLL:
    """libfringe_b200_prof.so: the stand-alone microbenchmarks (bench.py, scripts/)."""
    p = C.CDLL(PROF_LIB_PATH)
    for name in ("fringe_prof_fp32_peak", "fringe_prof_mma_tf32_rate", "fringe_prof_mma_f16_rate", "fringe_prof_mma_f16_k8_rate", "fringe_prof_fp64_peak"):
        getattr(p, name).argtypes = [C.c_int, C.POINTER(C.c_double)]
    p.fringe_prof_block_fma_rate.argtypes = [C.c_int, C.POINTER(C.c_double)]
    return p

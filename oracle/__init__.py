"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import
this package.  ``fringe_b200`` never does: the product path has no CPU fallback.

Two builds of the same loop text (``oracle/loops.hpp``) exist:

* ``load("port")``      -> ``oracle/liboracle.so``; per-pair / per-matrix workers restated in
  ``oracle/restated.hpp`` (each function cites the reference file:line it follows).
* ``load("reference")`` -> ``oracle/_ref/libfringe_ref.so``; workers are the reference's own
  ``KS2sample.hpp`` / ``AD2unique.hpp`` / ``ulongmask.hpp`` / ``EigenLapack.hpp`` compiled in
  place from ``/root/reference`` (built in the authoring container only; the ``.so`` travels).

``load()`` prefers the reference build when it is present.

``ref_driver(name)`` gives the reference's own block drivers (``src/<name>/<name>.cpp`` compiled unmodified against the
GDAL / Armadillo stand-ins of ``oracle/shims/``, ``oracle/ref_drivers/``): the restated loops are pinned to them in
``tests/test_reference_drivers_cpu.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(_HERE, "liboracle.so"),
    "reference": os.path.join(_HERE, "_ref", "libfringe_ref.so"),
}
KS2, AD2 = 0, 1
EVD, MLE, STBAS = 0, 1, 2
VARIANT_EVD, VARIANT_PHASE_LINK = 0, 1


def build(ref: bool | None = None) -> None:
    """Compile the oracle (and, where /root/reference exists, the reference-header build)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref is None:
        ref = os.path.isdir("/root/reference")
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
        subprocess.check_call(["make", "-s", "-C", _HERE, "refdrivers"])


def nulong(Nx: int, Ny: int) -> int:
    return int(np.ceil(((2 * Ny + 1) * (2 * Nx + 1)) / 32.0))


class Oracle:
    def __init__(self, path: str):
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # as src/evd/evd.py:43
        self.lib = lib = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        u8 = C.POINTER(C.c_uint8)
        u32 = C.POINTER(C.c_uint32)
        i32 = C.POINTER(C.c_int32)
        lib.oracle_kind.restype = C.c_char_p
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_set_threads.argtypes = [C.c_int]
        lib.oracle_ks2_prob.restype = C.c_double
        lib.oracle_ks2_prob.argtypes = [fp, fp, C.c_int]
        lib.oracle_kolmogorov_prob.restype = C.c_double
        lib.oracle_kolmogorov_prob.argtypes = [C.c_double]
        lib.oracle_ad2_prob.restype = C.c_double
        lib.oracle_ad2_prob.argtypes = [fp, fp, C.c_int]
        lib.oracle_ad2_sigma.restype = C.c_double
        lib.oracle_ad2_sigma.argtypes = [C.c_int]
        lib.oracle_ad2_pvalue_of_stat.restype = C.c_double
        lib.oracle_ad2_pvalue_of_stat.argtypes = [C.c_double, C.c_int]
        lib.oracle_mask_setbit.argtypes = [u32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_mask_getbit.argtypes = [u32, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_eig_extreme.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp]
        lib.oracle_pd_inverse.argtypes = [dp, C.c_int]
        lib.oracle_nmap_block.argtypes = [fp, u8, dp] + [C.c_int] * 6 + [C.c_double, i32, u32, fp]
        lib.oracle_evd_block.argtypes = [fp, u32] + [C.c_int] * 12 + [fp, fp, fp, i32]
        lib.oracle_ampdispersion_block.argtypes = [fp, dp, C.c_int, C.c_int, C.c_int, fp, fp]
        lib.oracle_despeck_block.argtypes = [fp, fp, u32] + [C.c_int] * 7 + [fp]
        lib.oracle_cmul.argtypes = [fp, fp, C.c_long, fp]
        lib.oracle_cmul.restype = None
        lib.oracle_calamp_block.argtypes = [fp, u8, C.c_long, C.c_int, dp, dp]
        lib.oracle_calamp_block.restype = None
        lib.oracle_integrate_ps.argtypes = [fp, fp, fp, fp, u8, C.c_long, fp]
        lib.oracle_integrate_ps.restype = None
        self.kind = lib.oracle_kind().decode()

    # -- helpers -------------------------------------------------------------------
    @staticmethod
    def _p(a, ty):
        return None if a is None else a.ctypes.data_as(C.POINTER(ty))

    def threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def set_threads(self, n: int) -> None:
        self.lib.oracle_set_threads(int(n))

    # -- single tests --------------------------------------------------------------
    def ks2_prob(self, a, b) -> float:
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        return float(self.lib.oracle_ks2_prob(self._p(a, C.c_float), self._p(b, C.c_float), a.size))

    def kolmogorov_prob(self, z: float) -> float:
        return float(self.lib.oracle_kolmogorov_prob(float(z)))

    def ad2_prob(self, a, b) -> float:
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        return float(self.lib.oracle_ad2_prob(self._p(a, C.c_float), self._p(b, C.c_float), a.size))

    def ad2_sigma(self, n: int) -> float:
        return float(self.lib.oracle_ad2_sigma(int(n)))

    def ad2_pvalue_of_stat(self, a2: float, n: int) -> float:
        return float(self.lib.oracle_ad2_pvalue_of_stat(float(a2), int(n)))

    def mask_setbit(self, words, Ny, Nx, dy, dx, on=True):
        self.lib.oracle_mask_setbit(self._p(words, C.c_uint32), Ny, Nx, dy, dx, int(on))

    def mask_getbit(self, words, Ny, Nx, dy, dx) -> bool:
        return bool(self.lib.oracle_mask_getbit(self._p(words, C.c_uint32), Ny, Nx, dy, dx))

    def eig_extreme(self, A, largest=True, want_vec=True):
        """A: (n,n) Hermitian complex128 (upper triangle is what LAPACK reads)."""
        A = np.array(A, dtype=np.complex128)
        n = A.shape[0]
        buf = np.asfortranarray(A).copy(order="F")
        val = np.zeros(1)
        vec = np.zeros(n, np.complex128)
        info = self.lib.oracle_eig_extreme(self._p(buf, C.c_double), n, int(largest), int(want_vec),
                                           self._p(val, C.c_double), self._p(vec, C.c_double))
        return info, float(val[0]), vec

    def pd_inverse(self, A):
        buf = np.asfortranarray(np.array(A, dtype=np.complex128)).copy(order="F")
        info = self.lib.oracle_pd_inverse(self._p(buf, C.c_double), buf.shape[0])
        return info, np.array(buf)

    # -- block drivers -------------------------------------------------------------
    def nmap_block(self, slc, Nx, Ny, method=KS2, thresh=0.05, mask=None, alpha=None,
                   want_amp=False):
        """slc: (bands, lines, cols) complex64 -> count (lines, cols) int32,
        wts (lines, cols, nulong) uint32 [, sorted amplitudes (lines, cols, bands)]."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        nu = nulong(Nx, Ny)
        count = np.zeros((lines, cols), np.int32)
        wts = np.zeros((lines, cols, nu), np.uint32)
        amp = np.zeros((lines, cols, bands), np.float32) if want_amp else None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, np.float64)
        rc = self.lib.oracle_nmap_block(self._p(slc, C.c_float), self._p(mask, C.c_uint8),
                                        self._p(alpha, C.c_double), cols, lines, bands, Nx, Ny,
                                        int(method), float(thresh), self._p(count, C.c_int32),
                                        self._p(wts, C.c_uint32), self._p(amp, C.c_float))
        if rc != 0:
            raise RuntimeError(f"oracle_nmap_block rc={rc}")
        return (count, wts, amp) if want_amp else (count, wts)

    def evd_block(self, slc, wts, Nx, Ny, method=EVD, bandwidth=-1, mini_stack_count=1,
                  variant=VARIANT_EVD, min_neighbors=2, first_line=0, n_lines=None,
                  want_npix=False):
        """slc: (bands, lines, cols) complex64; wts (lines, cols, nulong) uint32 ->
        out (bands, lines, cols) complex64, tcorr (lines, cols) f32, comp (lines, cols) c64."""
        slc = np.ascontiguousarray(slc, np.complex64)
        wts = np.ascontiguousarray(wts, np.uint32)
        bands, lines, cols = slc.shape
        if n_lines is None:
            n_lines = lines - first_line
        out = np.zeros((bands, lines, cols), np.complex64)
        tcorr = np.zeros((lines, cols), np.float32)
        comp = np.zeros((lines, cols), np.complex64)
        npix = np.zeros((lines, cols), np.int32) if want_npix else None
        rc = self.lib.oracle_evd_block(self._p(slc, C.c_float), self._p(wts, C.c_uint32), cols, lines,
                                       bands, Nx, Ny, first_line, n_lines, int(method), int(bandwidth),
                                       int(mini_stack_count), int(variant), int(min_neighbors),
                                       self._p(out, C.c_float), self._p(tcorr, C.c_float),
                                       self._p(comp, C.c_float), self._p(npix, C.c_int32))
        if rc != 0:
            raise RuntimeError(f"oracle_evd_block rc={rc}")
        return (out, tcorr, comp, npix) if want_npix else (out, tcorr, comp)

    def ampdispersion_block(self, slc, alpha=None):
        """slc (bands, lines, cols) complex64 -> amplitude dispersion, mean amplitude (float32 each)."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, np.float64)
        da = np.empty((lines, cols), np.float32)
        mean = np.empty((lines, cols), np.float32)
        rc = self.lib.oracle_ampdispersion_block(self._p(slc.view(np.float32), C.c_float), self._p(alpha, C.c_double),
                                                 cols, lines, bands, self._p(da, C.c_float), self._p(mean, C.c_float))
        assert rc == 0
        return da, mean

    def despeck_block(self, z1, wts, Nx, Ny, z2=None, coherence=False, first_line=0, n_lines=None):
        """SHP-weighted average (src/despeck/despeck.cpp): z1 [, z2] (lines, cols) complex64 -> complex64."""
        z1 = np.ascontiguousarray(z1, np.complex64)
        lines, cols = z1.shape
        if z2 is not None:
            z2 = np.ascontiguousarray(z2, np.complex64)
        wts = np.ascontiguousarray(wts, np.uint32)
        if n_lines is None:
            n_lines = lines - first_line
        out = np.zeros((lines, cols), np.complex64)
        rc = self.lib.oracle_despeck_block(self._p(z1.view(np.float32), C.c_float),
                                           None if z2 is None else self._p(z2.view(np.float32), C.c_float),
                                           self._p(wts, C.c_uint32), cols, lines, Nx, Ny, first_line, n_lines,
                                           1 if coherence else 0, self._p(out.view(np.float32), C.c_float))
        assert rc == 0
        return out

    def calamp_block(self, slc, mask=None):
        """Per band: sum of the valid amplitudes, number of valid pixels (src/calamp/calamp.cpp:207-226)."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands = slc.shape[0]
        npix = slc[0].size
        sums, counts = np.zeros(bands), np.zeros(bands)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self.lib.oracle_calamp_block(self._p(slc.view(np.float32), C.c_float), self._p(mask, C.c_uint8), npix, bands,
                                     self._p(sums, C.c_double), self._p(counts, C.c_double))
        return sums, counts

    def integrate_ps(self, ds_i, ds_j, slc_i, slc_j, ps):
        """python/integratePS.py:97-130 for one pair."""
        arrs = [np.ascontiguousarray(a, np.complex64) for a in (ds_i, ds_j, slc_i, slc_j)]
        ps = np.ascontiguousarray(ps, np.uint8)
        out = np.empty_like(arrs[0])
        self.lib.oracle_integrate_ps(*[self._p(a.view(np.float32), C.c_float) for a in arrs], self._p(ps, C.c_uint8), out.size,
                                     self._p(out.view(np.float32), C.c_float))
        return out

    def cmul(self, a, b):
        """Datum adjustment product a * b (complex64 in, double arithmetic, complex64 out)."""
        a = np.ascontiguousarray(a, np.complex64)
        b = np.ascontiguousarray(b, np.complex64)
        out = np.empty_like(a)
        self.lib.oracle_cmul(self._p(a.view(np.float32), C.c_float), self._p(b.view(np.float32), C.c_float),
                             a.size, self._p(out.view(np.float32), C.c_float))
        return out


_CACHE: dict[str, Oracle] = {}


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def load(kind: str | None = None) -> Oracle:
    if kind is None:
        kind = "reference" if available("reference") else "port"
    if kind not in _CACHE:
        if not available(kind):
            if kind == "port":
                build(ref=False)
            else:
                raise FileNotFoundError(_PATHS[kind] + " (build with `make -C oracle ref` where /root/reference exists)")
        _CACHE[kind] = Oracle(_PATHS[kind])
    return _CACHE[kind]


# ---- the reference's own block drivers (oracle/ref_drivers/, built by `make -C oracle refdrivers`) -----------------
REF_DRIVERS = ("nmap", "evd", "phase_link", "despeck", "ampdispersion", "calamp")
_REF_DRIVER_CACHE: dict[str, C.CDLL] = {}


def ref_driver_available(name: str, openmp: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{name}{'_omp' if openmp else ''}.so"))


def ref_driver(name: str, openmp: bool = False) -> C.CDLL:
    """ctypes handle of the reference driver `name` (src/<name>/<name>.cpp compiled unmodified against the GDAL /
    Armadillo stand-ins of oracle/shims/; file in, file out, like the reference's command line).  openmp=True: the
    multi-threaded build (nmap / evd / phase_link only), for timing -- nmap.cpp's pair loop races there."""
    key = name + ("_omp" if openmp else "")
    if key not in _REF_DRIVER_CACHE:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        lib = C.CDLL(os.path.join(_HERE, "_ref", f"libref_{key}.so"))
        s, i, d = C.c_char_p, C.c_int, C.c_double
        if name == "nmap":
            lib.ref_nmap.argtypes = [s, s, s, s, i, i, s, d, i, i]
        elif name in ("evd", "phase_link"):
            getattr(lib, "ref_" + name).argtypes = [s, s, s, s, s, i, i, s, i, i, i, i, i]
        elif name == "despeck":
            lib.ref_despeck.argtypes = [s, s, s, i, i, i, i, i, i, i]
        elif name == "ampdispersion":
            lib.ref_ampdispersion.argtypes = [s, s, s, i, i, i]
        elif name == "calamp":
            lib.ref_calamp.argtypes = [s, s, s, d, i, i, i, i, C.POINTER(C.c_double)]
        _REF_DRIVER_CACHE[key] = lib
    return _REF_DRIVER_CACHE[key]

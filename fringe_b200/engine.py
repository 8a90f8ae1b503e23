"""Thin Python face of the C ABI: one ``Context`` per GPU.

Host-array calls (``nmap_block`` / ``evd_block``) go through ``fringe_nmap_block`` /
``fringe_evd_block`` -- the same entry points the C++ block drivers use -- so the parity tests
exercise exactly the drop-in boundary.  The ``*_device`` calls take torch CUDA tensors (torch is
only the allocator and stream owner here) and queue work on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FringeError, lib

METHODS_NMAP = {"KS2": _lib.NMAP_KS2, "AD2": _lib.NMAP_AD2}
METHODS_EVD = {"EVD": _lib.EVD_EVD, "MLE": _lib.EVD_MLE, "STBAS": _lib.EVD_STBAS}


def nulong(Nx: int, Ny: int) -> int:
    return int(lib.fringe_nulong(Nx, Ny))


def device_count() -> int:
    n = C.c_int(0)
    lib.fringe_device_count(C.byref(n))
    return n.value


def ks2_critical_count(bands: int, pvalue: float):
    k, m = C.c_int(0), C.c_double(0)
    rc = lib.fringe_ks2_critical_count(bands, pvalue, C.byref(k), C.byref(m))
    if rc:
        raise FringeError(rc, lib.fringe_status_string(rc).decode())
    return k.value, m.value


def ad2_critical_sum(bands: int, pvalue: float) -> float:
    s = C.c_double(0)
    rc = lib.fringe_ad2_critical_sum(bands, pvalue, C.byref(s))
    if rc:
        raise FringeError(rc, lib.fringe_status_string(rc).decode())
    return s.value


def ad2_sigma(bands: int) -> float:
    s = C.c_double(0)
    rc = lib.fringe_ad2_sigma(bands, C.byref(s))
    if rc:
        raise FringeError(rc, lib.fringe_status_string(rc).decode())
    return s.value


def _method(table, m):
    if isinstance(m, str):
        if m.upper() not in table:
            return 99          # let the library report FRINGE_ERR_METHOD like nmap_process does
        return table[m.upper()]
    return int(m)


class Context:
    """Owns a ``fringe_ctx`` (stream + device workspaces) on one GPU."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.fringe_create(int(device), C.byref(h))
        if rc:
            raise FringeError(rc, lib.fringe_status_string(rc).decode())
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib.fringe_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise FringeError(rc, lib.fringe_last_error(self._h).decode() or lib.fringe_status_string(rc).decode())

    def synchronize(self):
        self._check(lib.fringe_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib.fringe_launch_count(self._h))

    def evd_stats(self) -> dict:
        arr = (C.c_int64 * 8)()
        self._check(lib.fringe_evd_stats(self._h, arr))
        return {"pixels": arr[0], "power_iterations": arr[1], "fp64_pixels": arr[2], "capped": arr[3],
                "factorisations": arr[4], "fp32_recomputed": arr[5]}

    def force_generic(self, on: bool) -> None:
        """Profiling / A-B tests only: send evd calls to the generic any-N kernel."""
        self._check(lib.fringe_prof_force_generic(self._h, 1 if on else 0))

    KERNELS = {"amp_sort": 0, "nmap": 1, "transpose": 2, "evd": 3, "cmul": 4, "despeck": 5, "ampdispersion": 6}

    def last_kernel_ms(self, kernel: str) -> float:
        ms = C.c_float(0)
        self._check(lib.fringe_last_kernel_ms(self._h, self.KERNELS[kernel], C.byref(ms)))
        return float(ms.value)

    def fp32_peak_tflops(self) -> float:
        t = C.c_double(0)
        self._check(_lib.prof_lib().fringe_prof_fp32_peak(self.device, C.byref(t)))
        return float(t.value)

    def fp64_peak_tflops(self) -> float:
        t = C.c_double(0)
        self._check(_lib.prof_lib().fringe_prof_fp64_peak(self.device, C.byref(t)))
        return float(t.value)

    # ---- host arrays -----------------------------------------------------------------------
    def nmap_block(self, slc, Nx, Ny, method="KS2", pvalue=0.05, mask=None, alpha=None):
        """slc (bands, lines, cols) complex64 -> count (lines, cols) int32, wts (lines, cols, nulong) uint32."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        nu = nulong(Nx, Ny)
        count = np.empty((lines, cols), np.int32)
        wts = np.empty((lines, cols, nu), np.uint32)
        mask_p = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mask_p = mask.ctypes.data
        alpha_p = None
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, np.float64)
            alpha_p = alpha.ctypes.data
        self._check(lib.fringe_nmap_block(self._h, slc.ctypes.data, mask_p, alpha_p, cols, lines, bands,
                                          Nx, Ny, _method(METHODS_NMAP, method), float(pvalue),
                                          count.ctypes.data, wts.ctypes.data))
        return count, wts

    def evd_block(self, slc, wts, Nx, Ny, method="EVD", bandwidth=-1, mini_stack_count=1,
                  variant=_lib.VARIANT_EVD, min_neighbors=2, first_line=0, n_lines=None):
        """-> out (bands, lines, cols) complex64, tcorr (lines, cols) f32, comp (lines, cols) c64;
        lines outside [first_line, first_line+n_lines) are returned as zeros."""
        slc = np.ascontiguousarray(slc, np.complex64)
        wts = np.ascontiguousarray(wts, np.uint32)
        bands, lines, cols = slc.shape
        if n_lines is None:
            n_lines = lines - first_line
        out = np.zeros((bands, lines, cols), np.complex64)
        tcorr = np.zeros((lines, cols), np.float32)
        comp = np.zeros((lines, cols), np.complex64)
        self._check(lib.fringe_evd_block(self._h, slc.ctypes.data, wts.ctypes.data, cols, lines, bands, Nx, Ny,
                                         first_line, n_lines, _method(METHODS_EVD, method), int(bandwidth),
                                         int(mini_stack_count), int(variant), int(min_neighbors),
                                         out.ctypes.data, tcorr.ctypes.data, comp.ctypes.data))
        return out, tcorr, comp

    def nmap_evd_block(self, slc, Nx, Ny, nmap_method="KS2", pvalue=0.05, mask=None, alpha=None,
                       method="EVD", bandwidth=-1, mini_stack_count=1, variant=_lib.VARIANT_EVD,
                       min_neighbors=2, first_line=0, n_lines=None, want_mask=True):
        """Both stages on one upload (``fringe_nmap_evd_block``).
        -> count, wts (None, None if not want_mask), out, tcorr, comp as nmap_block / evd_block."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        if n_lines is None:
            n_lines = lines - first_line
        nu = nulong(Nx, Ny)
        count = np.empty((lines, cols), np.int32) if want_mask else None
        wts = np.empty((lines, cols, nu), np.uint32) if want_mask else None
        mask_p = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mask_p = mask.ctypes.data
        alpha_p = None
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, np.float64)
            alpha_p = alpha.ctypes.data
        out = np.zeros((bands, lines, cols), np.complex64)
        tcorr = np.zeros((lines, cols), np.float32)
        comp = np.zeros((lines, cols), np.complex64)
        self._check(lib.fringe_nmap_evd_block(
            self._h, slc.ctypes.data, mask_p, alpha_p, cols, lines, bands, Nx, Ny,
            _method(METHODS_NMAP, nmap_method), float(pvalue), first_line, n_lines,
            _method(METHODS_EVD, method), int(bandwidth), int(mini_stack_count), int(variant),
            int(min_neighbors), count.ctypes.data if want_mask else None,
            wts.ctypes.data if want_mask else None, out.ctypes.data, tcorr.ctypes.data, comp.ctypes.data))
        return count, wts, out, tcorr, comp

    def sequential_block(self, slc, wts, Nx, Ny, mini_stack_size, method="MLE", bandwidth=-1, first_line=0, n_lines=None,
                         want_adjusted=True):
        """The whole ministack chain on one upload (``fringe_sequential_block``).
        slc (n_dates, lines, cols) complex64 -> dict(out_mini, tcorr_mini, comp, out_datum, tcorr_datum, adjusted)."""
        slc = np.ascontiguousarray(slc, np.complex64)
        wts = np.ascontiguousarray(wts, np.uint32)
        n_dates, lines, cols = slc.shape
        if n_lines is None:
            n_lines = lines - first_line
        nmini = -(-n_dates // mini_stack_size)
        res = {"out_mini": np.zeros((n_dates, lines, cols), np.complex64), "tcorr_mini": np.zeros((nmini, lines, cols), np.float32),
               "comp": np.zeros((nmini, lines, cols), np.complex64), "out_datum": np.zeros((nmini, lines, cols), np.complex64),
               "tcorr_datum": np.zeros((lines, cols), np.float32),
               "adjusted": np.zeros((n_dates, lines, cols), np.complex64) if want_adjusted else None}
        self._check(lib.fringe_sequential_block(
            self._h, slc.ctypes.data, wts.ctypes.data, cols, lines, n_dates, Nx, Ny, first_line, n_lines, int(mini_stack_size),
            _method(METHODS_EVD, method), int(bandwidth), res["out_mini"].ctypes.data, res["tcorr_mini"].ctypes.data,
            res["comp"].ctypes.data, res["out_datum"].ctypes.data, res["tcorr_datum"].ctypes.data,
            res["adjusted"].ctypes.data if want_adjusted else None))
        return res

    def ampdispersion_block(self, slc, alpha=None):
        """slc (bands, lines, cols) complex64 -> amplitude dispersion, mean amplitude (float32 each)."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        alpha_p = None
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, np.float64)
            alpha_p = alpha.ctypes.data
        da = np.empty((lines, cols), np.float32)
        mean = np.empty((lines, cols), np.float32)
        self._check(lib.fringe_ampdispersion_block(self._h, slc.ctypes.data, alpha_p, cols, lines, bands,
                                                   da.ctypes.data, mean.ctypes.data))
        return da, mean

    def ampdispersion_block_device(self, slc, alpha=None):
        import torch
        assert slc.is_cuda and slc.dtype == torch.complex64 and slc.is_contiguous()
        bands, lines, cols = slc.shape
        da = torch.empty((lines, cols), dtype=torch.float32, device=slc.device)
        mean = torch.empty_like(da)
        self._check(lib.fringe_ampdispersion_block_device(
            self._h, slc.data_ptr(), None if alpha is None else alpha.data_ptr(), cols, lines, bands,
            da.data_ptr(), mean.data_ptr(), self._stream()))
        return da, mean

    def calamp_block(self, slc, mask=None):
        """Sum of valid amplitudes and number of valid pixels per band (``fringe_calamp_block``)."""
        slc = np.ascontiguousarray(slc, np.complex64)
        bands, lines, cols = slc.shape
        sums, counts = np.zeros(bands), np.zeros(bands)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self._check(lib.fringe_calamp_block(self._h, slc.ctypes.data, None if mask is None else mask.ctypes.data, cols, lines,
                                            bands, sums.ctypes.data, counts.ctypes.data))
        return sums, counts

    def integrate_ps(self, ds_i, ds_j, slc_i, slc_j, ps):
        """Wrapped PS + DS interferogram of one pair (``fringe_integrate_ps``)."""
        arrs = [np.ascontiguousarray(a, np.complex64) for a in (ds_i, ds_j, slc_i, slc_j)]
        ps = np.ascontiguousarray(ps, np.uint8)
        out = np.empty_like(arrs[0])
        self._check(lib.fringe_integrate_ps(self._h, *[a.ctypes.data for a in arrs], ps.ctypes.data, out.size, out.ctypes.data))
        return out

    def ps_coherence(self, tcorr, ps, value=0.95):
        tcorr = np.ascontiguousarray(tcorr, np.float32)
        ps = np.ascontiguousarray(ps, np.uint8)
        out = np.empty_like(tcorr)
        self._check(lib.fringe_ps_coherence(self._h, tcorr.ctypes.data, ps.ctypes.data, out.size, float(value), out.ctypes.data))
        return out

    def despeck_block(self, z1, wts, Nx, Ny, z2=None, coherence=False, first_line=0, n_lines=None):
        """SHP-weighted average (``fringe_despeck_block``): z1 [, z2] (lines, cols) complex64 -> complex64."""
        z1 = np.ascontiguousarray(z1, np.complex64)
        lines, cols = z1.shape
        if z2 is not None:
            z2 = np.ascontiguousarray(z2, np.complex64)
            if z2.shape != z1.shape:
                raise ValueError("shape mismatch")
        wts = np.ascontiguousarray(wts, np.uint32)
        if n_lines is None:
            n_lines = lines - first_line
        out = np.zeros((lines, cols), np.complex64)
        self._check(lib.fringe_despeck_block(self._h, z1.ctypes.data, None if z2 is None else z2.ctypes.data,
                                             wts.ctypes.data, cols, lines, Nx, Ny, first_line, n_lines,
                                             1 if coherence else 0, out.ctypes.data))
        return out

    def despeck_block_device(self, z1, wts, Nx, Ny, z2=None, coherence=False, first_line=0, n_lines=None, out=None):
        import torch
        assert z1.is_cuda and z1.dtype == torch.complex64 and z1.is_contiguous() and wts.is_cuda and wts.is_contiguous()
        lines, cols = z1.shape
        if n_lines is None:
            n_lines = lines - first_line
        if out is None:
            out = torch.zeros_like(z1)
        self._check(lib.fringe_despeck_block_device(
            self._h, z1.data_ptr(), None if z2 is None else z2.data_ptr(), wts.data_ptr(), cols, lines, Nx, Ny,
            first_line, n_lines, 1 if coherence else 0, out.data_ptr(), self._stream()))
        return out

    def cmul(self, a, b):
        """Datum adjustment product a * b of two complex64 rasters (``fringe_cmul``)."""
        a = np.ascontiguousarray(a, np.complex64)
        b = np.ascontiguousarray(b, np.complex64)
        if a.shape != b.shape:
            raise ValueError("shape mismatch")
        out = np.empty_like(a)
        self._check(lib.fringe_cmul(self._h, a.ctypes.data, b.ctypes.data, a.size, out.ctypes.data))
        return out

    def cmul_device(self, a, b, out=None):
        import torch
        assert a.is_cuda and b.is_cuda and a.dtype == torch.complex64 and b.dtype == torch.complex64
        assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
        if out is None:
            out = torch.empty_like(a)
        self._check(lib.fringe_cmul_device(self._h, a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), self._stream()))
        return out

    # ---- device tensors (torch) ------------------------------------------------------------
    @staticmethod
    def _stream():
        import torch
        h = torch.cuda.current_stream().cuda_stream
        # torch's default stream has handle 0, which the C ABI reads as "use the context's own
        # stream"; name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so the
        # kernels really run on -- and CUDA events really time -- torch's current stream.
        return C.c_void_p(h if h else 1)

    def nmap_block_device(self, slc, Nx, Ny, method="KS2", pvalue=0.05, mask=None, alpha=None,
                          count=None, wts=None):
        """slc: torch complex64 CUDA tensor (bands, lines, cols), contiguous."""
        import torch
        assert slc.is_cuda and slc.dtype == torch.complex64 and slc.is_contiguous()
        bands, lines, cols = slc.shape
        nu = nulong(Nx, Ny)
        if count is None:
            count = torch.empty((lines, cols), dtype=torch.int32, device=slc.device)
        if wts is None:
            wts = torch.empty((lines, cols, nu), dtype=torch.int32, device=slc.device)
        self._check(lib.fringe_nmap_block_device(
            self._h, slc.data_ptr(), None if mask is None else mask.data_ptr(),
            None if alpha is None else alpha.data_ptr(), cols, lines, bands, Nx, Ny,
            _method(METHODS_NMAP, method), float(pvalue), count.data_ptr(), wts.data_ptr(), self._stream()))
        return count, wts

    def evd_block_device(self, slc, wts, Nx, Ny, method="EVD", bandwidth=-1, mini_stack_count=1,
                         variant=_lib.VARIANT_EVD, min_neighbors=2, first_line=0, n_lines=None,
                         out=None, tcorr=None, comp=None):
        import torch
        assert slc.is_cuda and slc.dtype == torch.complex64 and slc.is_contiguous()
        bands, lines, cols = slc.shape
        if n_lines is None:
            n_lines = lines - first_line
        if out is None:
            out = torch.zeros((bands, lines, cols), dtype=torch.complex64, device=slc.device)
        if tcorr is None:
            tcorr = torch.zeros((lines, cols), dtype=torch.float32, device=slc.device)
        if comp is None:
            comp = torch.zeros((lines, cols), dtype=torch.complex64, device=slc.device)
        self._check(lib.fringe_evd_block_device(
            self._h, slc.data_ptr(), wts.data_ptr(), cols, lines, bands, Nx, Ny, first_line, n_lines,
            _method(METHODS_EVD, method), int(bandwidth), int(mini_stack_count), int(variant),
            int(min_neighbors), out.data_ptr(), tcorr.data_ptr(), comp.data_ptr(), self._stream()))
        return out, tcorr, comp

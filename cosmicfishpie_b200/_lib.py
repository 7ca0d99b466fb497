"""ctypes binding of libcfp_b200.so (the C ABI declared in include/cfp_b200.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is usable, every
compute entry point raises.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C cosmicfishpie_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfp_b200.so")

# mirror of include/cfp_b200.h -----------------------------------------------------------------
ABI_VERSION = 1
BANK_DESC = 5
TAB_PLIN, TAB_PNL, TAB_F, TAB_D, TAB_PLIN_CB, TAB_PNL_CB, TAB_F_CB, NTAB = 0, 1, 2, 3, 4, 5, 6, 7
(SP_Z, SP_COSMO, SP_QPAR, SP_QPER, SP_HUBBLE, SP_DZ, SP_BIAS, SP_A1, SP_A2, SP_SIGMAP, SP_I0, SP_I1, SP_I2,
 SP_SVRESCALE, SP_FS8, SP_INVS8SQ, SP_KHFAC, SP_PSHOT, SP_NSCAL) = range(19)
SPF_AP, SPF_FOG, SPF_LINEAR, SPF_KH_BEFORE_ERR = 1, 2, 4, 8
MAT_LNP, MAT_DERIVS = 1, 2

EXPORTS = [
    "cfp_abi_version", "cfp_last_error", "cfp_ctx_create", "cfp_ctx_destroy", "cfp_ctx_sync",
    "cfp_ctx_launch_count", "cfp_ctx_last_run_ms", "cfp_peak_fp64", "cfp_spectro_plan_kernel_ms",
    "cfp_photo_plan_kernel_ms", "cfp_bank_upload", "cfp_bank_free", "cfp_bank_eval",
    "cfp_spectro_plan_create", "cfp_spectro_plan_destroy", "cfp_spectro_plan_set_slice", "cfp_spectro_plan_run",
    "cfp_spectro_plan_fisher_d", "cfp_spectro_plan_fetch", "cfp_photo_plan_create", "cfp_photo_plan_destroy",
    "cfp_photo_plan_set_slice", "cfp_photo_plan_run", "cfp_photo_plan_fisher_d", "cfp_photo_plan_fetch",
    "cfp_photo_plan_chi2", "cfp_spectro_plan_chi2", "cfp_cells_chi2", "cfp_wedge_chi2",
    "cfp_legendre_project", "cfp_legendre_covariance", "cfp_legendre_chi2",
    "cfp_bank_upload_rows", "cfp_host_alloc", "cfp_host_free", "cfp_spectro_plan_chi2_dev", "cfp_dev_upload",
    "cfp_dev_release", "cfp_photo_plan_create_zmap", "cfp_ctx_aux_stream", "cfp_ctx_aux_stream2", "cfp_photo_plan_create_ex",
    "cfp_photo_plan_set_gc_table", "cfp_spectro_plan_chi2_legendre", "cfp_spectro_plan_kernel_ms_part",
    "cfp_math_selftest", "cfp_spectro_plan_terms", "cfp_photo_plan_fetch_windows", "cfp_photo_fisher",
    "cfp_fisher_batch_post", "cfp_fisher_batch_add", "cfp_bank_combine", "cfp_bank_combine_rows", "cfp_photo_plan_create_zmap_shared",
    "cfp_photo_plan_chi2_launch", "cfp_photo_plan_chi2_collect", "cfp_bank_fit", "cfp_background_fit",
    "cfp_bank_eval_many", "cfp_bank_nowiggle", "cfp_bank_theta_moments", "cfp_photo_plan_set_ia_enla",
]


class CfpError(RuntimeError):
    pass


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_vp = C.c_void_p


def load():
    """Load the shared library (once).  Raises CfpError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CfpError(
            f"{LIB_PATH} not found: the CUDA extension is not built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C cosmicfishpie_b200/csrc`). "
            "cosmicfishpie_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.cfp_abi_version.restype = C.c_int
    lib.cfp_last_error.restype = C.c_char_p
    lib.cfp_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
    lib.cfp_ctx_destroy.argtypes = [_vp]
    lib.cfp_ctx_sync.argtypes = [_vp]
    lib.cfp_ctx_aux_stream.argtypes = [_vp]
    lib.cfp_ctx_aux_stream.restype = _vp
    lib.cfp_ctx_aux_stream2.argtypes = [_vp]
    lib.cfp_ctx_aux_stream2.restype = _vp
    lib.cfp_ctx_launch_count.argtypes = [_vp]
    lib.cfp_ctx_launch_count.restype = C.c_int64
    lib.cfp_ctx_last_run_ms.argtypes = [_vp]
    lib.cfp_ctx_last_run_ms.restype = C.c_double
    lib.cfp_peak_fp64.argtypes = [_vp, C.c_int, _dp]
    lib.cfp_spectro_plan_kernel_ms.argtypes = [_vp]
    lib.cfp_spectro_plan_kernel_ms.restype = C.c_double
    lib.cfp_spectro_plan_kernel_ms_part.argtypes = [_vp, C.c_int]
    lib.cfp_spectro_plan_kernel_ms_part.restype = C.c_double
    lib.cfp_spectro_plan_terms.argtypes = [_vp, C.c_int, _dp]
    lib.cfp_photo_plan_fetch_windows.argtypes = [_vp, _dp, _dp]
    lib.cfp_photo_fisher.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    lib.cfp_fisher_batch_post.argtypes = [_vp, C.c_int, C.c_int, _dp, C.c_int, _ip, C.c_double, _dp, _dp, _dp]
    lib.cfp_fisher_batch_add.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
    lib.cfp_bank_combine.argtypes = [_vp, _vp, C.c_int, _dp, C.POINTER(_vp)]
    lib.cfp_bank_combine_rows.argtypes = [_vp, _vp, C.c_int, _dp, C.POINTER(_vp)]
    lib.cfp_math_selftest.argtypes = [_vp, C.c_int, _dp, C.c_int, _dp]
    lib.cfp_math_selftest.restype = C.c_int
    lib.cfp_photo_plan_kernel_ms.argtypes = [_vp, C.c_int]
    lib.cfp_photo_plan_kernel_ms.restype = C.c_double
    lib.cfp_bank_upload.argtypes = [_vp, _dp, C.c_size_t, _lp, C.c_int, C.c_int, C.POINTER(_vp)]
    lib.cfp_bank_free.argtypes = [_vp, _vp]
    lib.cfp_bank_upload_rows.argtypes = [_vp, C.POINTER(_vp), _lp, _lp, C.c_int, C.c_int, C.POINTER(_vp)]
    lib.cfp_host_alloc.argtypes = [_vp, C.c_size_t, C.POINTER(_vp)]
    lib.cfp_host_free.argtypes = [_vp, _vp]
    lib.cfp_spectro_plan_chi2_dev.argtypes = [_vp, _vp, _dp, _dp, _vp]
    lib.cfp_spectro_plan_chi2_legendre.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp, _vp]
    lib.cfp_dev_upload.argtypes = [_vp, _dp, C.c_size_t, C.POINTER(_vp)]
    lib.cfp_dev_release.argtypes = [_vp, _vp]
    lib.cfp_bank_eval.argtypes = [_vp, _vp, C.c_int, C.c_int, _dp, _dp, C.c_size_t, _dp]
    lib.cfp_spectro_plan_create.argtypes = [
        _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, C.c_int, _dp, _dp,
        _dp, _dp, _ip, C.c_int, _dp, _dp, C.c_uint, C.c_int, C.c_int, C.POINTER(_vp)]
    lib.cfp_spectro_plan_destroy.argtypes = [_vp]
    lib.cfp_spectro_plan_set_slice.argtypes = [_vp, C.c_int64, C.c_int64]
    lib.cfp_spectro_plan_run.argtypes = [_vp, C.c_int, _vp]
    lib.cfp_spectro_plan_fisher_d.argtypes = [_vp]
    lib.cfp_spectro_plan_fisher_d.restype = _vp
    lib.cfp_spectro_plan_fetch.argtypes = [_vp, _dp, _dp, _dp, _dp]
    lib.cfp_photo_plan_create.argtypes = [
        _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, C.c_double, _dp, _dp, _dp,
        _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.POINTER(_vp)]
    lib.cfp_photo_plan_create_zmap.argtypes = [
        _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, C.c_double, _dp, _dp, _dp,
        _ip, _dp, _dp, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.POINTER(_vp)]
    lib.cfp_photo_plan_create_zmap_shared.argtypes = lib.cfp_photo_plan_create_zmap.argtypes
    lib.cfp_photo_plan_create_ex.argtypes = [
        _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, C.c_double, _dp, _dp, _dp,
        _ip, C.c_int, _ip, _dp, _dp, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int,
        C.POINTER(_vp)]
    lib.cfp_photo_plan_set_gc_table.argtypes = [_vp, C.c_int]
    lib.cfp_photo_plan_destroy.argtypes = [_vp]
    lib.cfp_photo_plan_set_slice.argtypes = [_vp, C.c_int, C.c_int]
    lib.cfp_photo_plan_run.argtypes = [_vp, _vp]
    lib.cfp_photo_plan_fisher_d.argtypes = [_vp]
    lib.cfp_photo_plan_fisher_d.restype = _vp
    lib.cfp_photo_plan_fetch.argtypes = [_vp, _dp, _dp, _dp]
    lib.cfp_photo_plan_chi2.argtypes = [_vp, _dp, C.c_int, _ip, _ip, _dp, _dp, _vp]
    lib.cfp_photo_plan_set_ia_enla.argtypes = [_vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]
    lib.cfp_bank_fit.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.POINTER(_vp), _dp, C.POINTER(_vp)]
    lib.cfp_background_fit.argtypes = [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, _dp]
    lib.cfp_bank_eval_many.argtypes = [_vp, _vp, C.c_int, C.c_size_t, _dp, _dp, _dp]
    lib.cfp_bank_nowiggle.argtypes = [_vp, _vp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _ip, _ip, _lp, _dp,
                                      C.c_size_t, _dp]
    lib.cfp_bank_theta_moments.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp]
    lib.cfp_photo_plan_chi2_launch.argtypes = [_vp, _dp, C.c_int, _ip, _ip, _dp, _vp]
    lib.cfp_photo_plan_chi2_collect.argtypes = [_vp, _dp]
    lib.cfp_spectro_plan_chi2.argtypes = [_vp, _dp, _dp, _dp, _dp, _vp]
    lib.cfp_cells_chi2.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp]
    lib.cfp_wedge_chi2.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.cfp_legendre_project.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    lib.cfp_legendre_covariance.argtypes = [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.cfp_legendre_chi2.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("cfp_abi_version",):
            pass
    if lib.cfp_abi_version() != ABI_VERSION:
        raise CfpError(f"ABI mismatch: library {lib.cfp_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().cfp_last_error().decode(errors="replace")
        raise CfpError(f"{what} failed (status {rc}): {msg}")


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: Optional[np.ndarray], typ=_dp):
    return None if a is None else a.ctypes.data_as(typ)


class Context:
    """One per GPU (cfp_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.handle = _vp()
        _check(self.lib.cfp_ctx_create(int(device), C.byref(self.handle)), "cfp_ctx_create")
        self.device = int(device)

    def sync(self):
        _check(self.lib.cfp_ctx_sync(self.handle), "cfp_ctx_sync")

    @property
    def aux_stream(self) -> int:
        """Handle of the context's second stream (pass as `stream=`)."""
        return int(self.lib.cfp_ctx_aux_stream(self.handle) or 0)

    @property
    def aux_stream2(self) -> int:
        """Handle of the context's third stream: alternate with `aux_stream` to overlap consecutive batches."""
        return int(self.lib.cfp_ctx_aux_stream2(self.handle) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.cfp_ctx_launch_count(self.handle))

    @property
    def last_run_ms(self) -> float:
        return float(self.lib.cfp_ctx_last_run_ms(self.handle))

    def math_selftest(self, which: int, x) -> np.ndarray:
        """f(x) on the device for the lattice loops' FP64 functions (cfp_math_selftest)."""
        x = f64(x).ravel()
        y = np.empty_like(x)
        _check(self.lib.cfp_math_selftest(self.handle, int(which), _p(x), x.size, _p(y)), "cfp_math_selftest")
        return y

    def photo_fisher(self, cov, dC, ell) -> np.ndarray:
        """cov [Nl][Nb][Nb], dC [npar][Nl][Nb][Nb], ell [Nl] -> F [npar][npar]  (cfp_photo_fisher)."""
        cov, dC, ell = f64(cov), f64(dC), f64(ell)
        npar, Nl, Nb = dC.shape[0], dC.shape[1], dC.shape[2]
        if cov.shape != (Nl, Nb, Nb) or dC.shape[3] != Nb or ell.shape != (Nl,):
            raise ValueError("photo_fisher: expected cov (Nl, Nb, Nb), dC (npar, Nl, Nb, Nb), ell (Nl,)")
        out = np.empty((npar, npar))
        _check(self.lib.cfp_photo_fisher(self.handle, Nl, Nb, npar, _p(cov), _p(dC), _p(ell), _p(out)), "cfp_photo_fisher")
        return out

    def fisher_batch_post(self, F, keep=None, coef=1.0):
        """F [B][n][n] -> (sigma [B][n], marginal [B][m][m] or None, fom [B] or None)  (cfp_fisher_batch_post)."""
        F = f64(F)
        B, n = F.shape[0], F.shape[1]
        keep = None if keep is None else i32(keep)
        m = 0 if keep is None else keep.size
        sigma = np.empty((B, n))
        marg = np.empty((B, m, m)) if m else None
        fom = np.empty(B) if m else None
        _check(self.lib.cfp_fisher_batch_post(self.handle, B, n, _p(F), m, _p(keep, _ip), float(coef), _p(sigma),
                                              _p(marg), _p(fom)), "cfp_fisher_batch_post")
        return sigma, marg, fom

    def fisher_batch_add(self, a, b, map_a, map_b, n_out) -> np.ndarray:
        a, b = f64(a), f64(b)
        out = np.empty((a.shape[0], n_out, n_out))
        _check(self.lib.cfp_fisher_batch_add(self.handle, a.shape[0], a.shape[1], b.shape[1], int(n_out),
                                             _p(i32(map_a), _ip), _p(i32(map_b), _ip), _p(a), _p(b), _p(out)),
               "cfp_fisher_batch_add")
        return out

    def peak_fp64(self, kind: int) -> float:
        """Measured FP64 TFLOP/s: kind 0 = DFMA, 1 = DMMA."""
        out = C.c_double(0.0)
        _check(self.lib.cfp_peak_fp64(self.handle, int(kind), C.byref(out)), "cfp_peak_fp64")
        return float(out.value)

    def cells_chi2(self, theory, data, blk_off, blk_n, blk_w) -> np.ndarray:
        """theory [S][Nl][Nb][Nb], data [Nl][Nb][Nb] (noise included) -> chi2[S]  (cfp_cells_chi2)."""
        theory, data = f64(theory), f64(data)
        S, Nl, Nb = theory.shape[0], theory.shape[1], theory.shape[2]
        off, n, w = i32(blk_off), i32(blk_n), f64(blk_w)
        out = np.empty(S)
        _check(self.lib.cfp_cells_chi2(self.handle, S, Nl, Nb, _p(theory), _p(data), off.size, _p(off, _ip), _p(n, _ip),
                                       _p(w), _p(out)), "cfp_cells_chi2")
        return out

    def wedge_chi2(self, theory, data, k, wk, wmu, vol) -> np.ndarray:
        """theory [S][Z][Nmu][Nk], data [Z][Nmu][Nk] -> chi2[S]  (cfp_wedge_chi2)."""
        theory, data = f64(theory), f64(data)
        S, Z, Nmu, Nk = theory.shape
        out = np.empty(S)
        _check(self.lib.cfp_wedge_chi2(self.handle, S, Z, Nmu, Nk, _p(theory), _p(data), _p(f64(k)), _p(f64(wk)),
                                       _p(f64(wmu)), _p(f64(vol)), _p(out)), "cfp_wedge_chi2")
        return out

    def legendre_project(self, lat, lw) -> np.ndarray:
        """lat [S][Z][Nmu][Nk], lw [3][Nmu] -> P_ell [S][Nk][Z][3]  (cfp_legendre_project)."""
        lat, lw = f64(lat), f64(lw)
        S, Z, Nmu, Nk = lat.shape
        out = np.empty((S, Nk, Z, 3))
        _check(self.lib.cfp_legendre_project(self.handle, S, Z, Nmu, Nk, _p(lat), _p(lw), _p(out)), "cfp_legendre_project")
        return out

    def legendre_covariance(self, pl, k, vol, m):
        """pl [Nk][Z][3] -> (cov, icov) [Nk][Z][3][3]  (cfp_legendre_covariance)."""
        pl = f64(pl)
        Nk, Z = pl.shape[:2]
        cov, icov = np.empty((Nk, Z, 3, 3)), np.empty((Nk, Z, 3, 3))
        _check(self.lib.cfp_legendre_covariance(self.handle, Nk, Z, _p(pl), _p(f64(k)), _p(f64(vol)), _p(f64(m)), _p(cov),
                                                _p(icov)), "cfp_legendre_covariance")
        return cov, icov

    def legendre_chi2(self, pl_data, pl_theory, icov) -> np.ndarray:
        pl_data, pl_theory, icov = f64(pl_data), f64(pl_theory), f64(icov)
        S, Nk, Z, nl = pl_theory.shape
        if nl > 3 or pl_data.shape != (Nk, Z, nl) or icov.shape != (Nk, Z, nl, nl):
            raise ValueError("legendre_chi2: expected P_ell (Nk, Nz, n<=3) and an inverse covariance (Nk, Nz, n, n)")
        if nl < 3:  # the kernel contracts three multipoles: pad with zeros (contributes nothing)
            pd, pt, ic = np.zeros((Nk, Z, 3)), np.zeros((S, Nk, Z, 3)), np.zeros((Nk, Z, 3, 3))
            pd[..., :nl], pt[..., :nl], ic[..., :nl, :nl] = pl_data, pl_theory, icov
            pl_data, pl_theory, icov = pd, pt, ic
        out = np.empty(S)
        _check(self.lib.cfp_legendre_chi2(self.handle, S, Nk, Z, _p(pl_data), _p(pl_theory), _p(icov), _p(out)),
               "cfp_legendre_chi2")
        return out

    def close(self):
        if self.handle:
            self.lib.cfp_ctx_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide context for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


_PINNED_POOL = {}  # n_doubles -> [address, ...]: page-locking is slow (ms per call), so released blocks are kept
_PINNED_POOL_MAX_BYTES = 1 << 30
_pinned_pool_bytes = 0


class PinnedBuffer:
    """Page-locked host array of float64 (cfp_host_alloc) -- buffers that are uploaded again and again.
    close() returns the block to a process-wide pool keyed by size (table rows of one bank all have one size)."""

    def __init__(self, ctx: Context, n_doubles: int):
        global _pinned_pool_bytes
        self.ctx, self.n = ctx, int(n_doubles)
        free = _PINNED_POOL.get(self.n)
        if free:
            self.ptr = _vp(free.pop())
            _pinned_pool_bytes -= self.n * 8
        else:
            self.ptr = _vp()
            _check(ctx.lib.cfp_host_alloc(ctx.handle, self.n * 8, C.byref(self.ptr)), "cfp_host_alloc")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _dp), shape=(self.n,))

    def close(self):
        global _pinned_pool_bytes
        if self.ptr:
            self.array = None
            if _pinned_pool_bytes + self.n * 8 <= _PINNED_POOL_MAX_BYTES:
                _PINNED_POOL.setdefault(self.n, []).append(self.ptr.value)
                _pinned_pool_bytes += self.n * 8
            else:
                self.ctx.lib.cfp_host_free(None, self.ptr)
            self.ptr = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass


class DeviceArray:
    """float64 array resident on the device (cfp_dev_upload), e.g. the data vector of a likelihood."""

    def __init__(self, ctx: Context, host):
        host = f64(host)
        self.ctx, self.shape, self.ptr = ctx, host.shape, _vp()
        _check(ctx.lib.cfp_dev_upload(ctx.handle, _p(host), host.size, C.byref(self.ptr)), "cfp_dev_upload")

    def close(self):
        if self.ptr:
            if self.ctx.handle:  # a destroyed context has already returned its whole arena
                self.ctx.lib.cfp_dev_release(self.ctx.handle, self.ptr)
            self.ptr = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass


class BankRow:
    """The fitted tables of ONE cosmology packed for upload: [tx | ty | c] per slot in a pinned block plus the
    row-local descriptor.  Built once per cosmology and table selection, reused by every bank that needs it."""

    def __init__(self, ctx: Context, row):
        desc = np.zeros((NTAB, BANK_DESC), dtype=np.int64)
        parts, off = [], 0
        for slot in range(NTAB):
            t = row[slot] if slot < len(row) else None
            if t is None:
                continue
            tx, ty, cc = (f64(x).ravel() for x in t)
            assert cc.size == (tx.size - 4) * (ty.size - 4)
            desc[slot] = (tx.size, ty.size, off, off + tx.size, off + tx.size + ty.size)
            parts += [tx, ty, cc]
            off += tx.size + ty.size + cc.size
        self.buf = PinnedBuffer(ctx, off)
        np.concatenate(parts, out=self.buf.array)
        self.desc, self.n = desc, off


class Bank:
    """Device-resident cosmology table bank (cfp_bank): bicubic (tx, ty, c) per cosmology and table."""

    @classmethod
    def from_rows(cls, ctx: Context, rows):
        """Bank of pre-packed `BankRow`s (one per cosmology): no re-packing, page-locked H2D."""
        self = cls.__new__(cls)
        self.ctx, self.rows = ctx, list(rows)
        n = len(self.rows)
        self.desc = np.ascontiguousarray(np.stack([r.desc for r in self.rows]))
        ptrs = (_vp * n)(*[r.buf.ptr.value for r in self.rows])
        lens = np.array([r.n for r in self.rows], dtype=np.int64)
        self.blob = None
        self.n_cosmo = n
        self.nbytes = int(lens.sum()) * 8 + self.desc.nbytes
        self.handle = _vp()
        _check(ctx.lib.cfp_bank_upload_rows(ctx.handle, ptrs, _p(lens, _lp), _p(self.desc, _lp), n, NTAB,
                                            C.byref(self.handle)), "cfp_bank_upload_rows")
        return self

    def __init__(self, ctx: Context, tcks):
        """tcks[c][slot] = (tx, ty, c_flat) or None."""
        self.ctx = ctx
        n_cosmo = len(tcks)
        desc = np.zeros((n_cosmo, NTAB, BANK_DESC), dtype=np.int64)
        parts, off = [], 0
        for ci, row in enumerate(tcks):
            for slot in range(NTAB):
                t = row[slot] if slot < len(row) else None
                if t is None:
                    continue
                tx, ty, cc = (f64(x).ravel() for x in t)
                assert cc.size == (tx.size - 4) * (ty.size - 4)
                desc[ci, slot] = (tx.size, ty.size, off, off + tx.size, off + tx.size + ty.size)
                parts += [tx, ty, cc]
                off += tx.size + ty.size + cc.size
        self.blob = np.concatenate(parts)
        self.desc = desc
        self.n_cosmo = n_cosmo
        self.nbytes = self.blob.nbytes + desc.nbytes
        self.handle = _vp()
        _check(ctx.lib.cfp_bank_upload(ctx.handle, _p(self.blob), self.blob.size, _p(desc, _lp), n_cosmo, NTAB,
                                       C.byref(self.handle)), "cfp_bank_upload")

    @classmethod
    def fit(cls, ctx: Context, z, k, tables, scales) -> "Bank":
        """Bank fitted ON THE DEVICE from raw tables (cfp_bank_fit): z[Nz], k[NC][Nk], tables[c][slot] = [Nz][Nk] array or
        None (the same slots for every cosmology), scales[c][slot]."""
        z, k = f64(z), f64(k)
        NC, Nk = k.shape
        keep, ptrs = [], (_vp * (NC * NTAB))()
        sc = np.ones((NC, NTAB))
        for c, row in enumerate(tables):
            for slot in range(NTAB):
                t = row[slot] if slot < len(row) else None
                if t is None:
                    continue
                a = f64(t)
                assert a.shape == (z.size, Nk), (a.shape, z.size, Nk)
                keep.append(a)
                ptrs[c * NTAB + slot] = a.ctypes.data
                sc[c, slot] = scales[c][slot]
        self = cls.__new__(cls)
        self.ctx, self.rows, self.blob, self.desc, self.n_cosmo = ctx, [], None, None, NC
        self.nbytes = sum(a.nbytes for a in keep) + z.nbytes + k.nbytes
        self.handle = _vp()
        _check(ctx.lib.cfp_bank_fit(ctx.handle, NC, NTAB, z.size, Nk, _p(z), _p(k), ptrs, _p(sc), C.byref(self.handle)),
               "cfp_bank_fit")
        return self

    def eval_many(self, table: int, z, k) -> np.ndarray:
        """[n_cosmo][n] values of one table of EVERY cosmology at the points (z, k) (cfp_bank_eval_many)."""
        z, k = np.broadcast_arrays(f64(z), f64(k))
        z, k = f64(z).ravel(), f64(k).ravel()
        out = np.empty((self.n_cosmo, z.size))
        _check(self.ctx.lib.cfp_bank_eval_many(self.ctx.handle, self.handle, int(table), z.size, _p(z), _p(k), _p(out)),
               "cfp_bank_eval_many")
        return out

    def nowiggle(self, table: int, zq, ks, filt_of_c, filters) -> np.ndarray:
        """[n_cosmo][nzq][nsamp] no-wiggle spectra (cfp_bank_nowiggle); ks[n_cosmo][nsamp] sample wavenumbers,
        filters = [(taps[win], off, EL[ne][wc], ER[ne][wc]), ...], filt_of_c[n_cosmo] picks one per cosmology."""
        zq, ks = f64(zq).ravel(), f64(ks)
        assert ks.shape[0] == self.n_cosmo
        fdesc = np.zeros((len(filters), 4), dtype=np.int32)
        foff = np.zeros(len(filters), dtype=np.int64)
        parts, off = [], 0
        for i, (taps, o, EL, ER) in enumerate(filters):
            taps, EL, ER = f64(taps).ravel(), f64(EL), f64(ER)
            assert EL.shape == ER.shape
            fdesc[i] = (taps.size, o, EL.shape[0], EL.shape[1])
            foff[i] = off
            parts += [taps, EL.ravel(), ER.ravel()]
            off += taps.size + 2 * EL.size
        blob = np.concatenate(parts)
        out = np.empty((self.n_cosmo, zq.size, ks.shape[1]))
        fc = i32(filt_of_c)
        _check(self.ctx.lib.cfp_bank_nowiggle(self.ctx.handle, self.handle, int(table), zq.size, _p(zq), ks.shape[1], _p(ks),
                                              len(filters), _p(fc, _ip), _p(fdesc, _ip), _p(foff, _lp), _p(blob), blob.size,
                                              _p(out)), "cfp_bank_nowiggle")
        return out

    def theta_moments(self, tab_p: int, tab_f: int, zq, k) -> np.ndarray:
        """[n_cosmo][nzq][3] velocity moments (cfp_bank_theta_moments)."""
        zq, k = f64(zq).ravel(), f64(k).ravel()
        out = np.empty((self.n_cosmo, zq.size, 3))
        _check(self.ctx.lib.cfp_bank_theta_moments(self.ctx.handle, self.handle, int(tab_p), int(tab_f), zq.size, _p(zq),
                                                   k.size, _p(k), _p(out)), "cfp_bank_theta_moments")
        return out

    def combined(self, W, keep_base: bool = True) -> "Bank":
        """A new bank: this bank's rows followed by len(W) rows combined on the device, row m = sum_c W[m][c] * row_c
        (cfp_bank_combine): emulated cosmologies without host fits.  keep_base=False: the combined rows only
        (cfp_bank_combine_rows), in the order of W's rows."""
        W = f64(W)
        assert W.ndim == 2 and W.shape[1] == self.n_cosmo
        out = Bank.__new__(Bank)
        out.ctx, out.rows, out.blob, out.desc = self.ctx, [], None, None
        out.n_cosmo = (self.n_cosmo if keep_base else 0) + W.shape[0]
        out.nbytes = W.nbytes
        out.handle = _vp()
        fn = self.ctx.lib.cfp_bank_combine if keep_base else self.ctx.lib.cfp_bank_combine_rows
        _check(fn(self.ctx.handle, self.handle, W.shape[0], _p(W), C.byref(out.handle)), "cfp_bank_combine")
        return out

    def eval(self, cosmo: int, table: int, z, k) -> np.ndarray:
        z, k = np.broadcast_arrays(f64(z), f64(k))
        z, k = f64(z).ravel(), f64(k).ravel()
        out = np.empty_like(z)
        _check(self.ctx.lib.cfp_bank_eval(self.ctx.handle, self.handle, cosmo, table, _p(z), _p(k), z.size, _p(out)),
               "cfp_bank_eval")
        return out

    def close(self):
        if self.handle:
            if self.ctx.handle:  # a destroyed context has already returned its whole arena
                self.ctx.lib.cfp_bank_free(self.ctx.handle, self.handle)
            self.handle = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass


def background_fit(ctx: Context, z, hub, extra, zq=None, ntrap: int = 100, coef: bool = False):
    """Background functions of NC cosmologies fitted on the device (cfp_background_fit): rows H, comoving distance,
    angular-diameter distance, then the extra functions of z.  Returns the values [NC][3 + n_extra][nq] at the
    redshifts zq (None without zq) and, with coef=True, the B-spline coefficients [NC][3 + n_extra][Nz] as well."""
    z, hub = f64(z).ravel(), f64(hub)
    zq = f64(zq).ravel() if zq is not None else np.zeros(0)
    NC = hub.shape[0]
    extra = f64(extra) if extra is not None and np.size(extra) else np.zeros((NC, 0, z.size))
    assert hub.shape == (NC, z.size) and extra.shape[0] == NC and extra.shape[2] == z.size
    out = np.empty((NC, 3 + extra.shape[1], zq.size))
    cf = np.empty((NC, 3 + extra.shape[1], z.size)) if coef else None
    _check(ctx.lib.cfp_background_fit(ctx.handle, NC, z.size, _p(z), _p(hub), extra.shape[1], _p(extra), int(ntrap), zq.size,
                                      _p(zq), _p(out), _p(cf) if coef else None), "cfp_background_fit")
    out = out if zq.size else None
    return (out, cf) if coef else out


class SpectroPlan:
    def __init__(self, ctx: Context, bank: Bank, *, k, mu, wk, wmu, scal, alias, pnw_k, pnw_p, stw, stden, ps_bin,
                 ps_src, nbar, vol, flags, tab_plin=TAB_PLIN, tab_f=TAB_F):
        self.ctx, self.bank = ctx, bank
        self.k, self.mu, self.wk, self.wmu = f64(k), f64(mu), f64(wk), f64(wmu)
        self.scal = f64(scal)
        self.S, self.Z = self.scal.shape[:2]
        assert self.scal.shape[2] == SP_NSCAL
        self.alias = i32(alias)
        self.pnw_k, self.pnw_p = f64(pnw_k), f64(pnw_p)
        self.stw, self.stden, self.ps_bin = f64(stw), f64(stden), i32(ps_bin)
        self.npar = self.stw.shape[0]
        self.nbar, self.vol = f64(nbar), f64(vol)
        self.Nk, self.Nmu = self.k.size, self.mu.size
        nnw = self.pnw_k.shape[-1]
        assert self.pnw_k.shape == (bank.n_cosmo, nnw) and self.pnw_p.shape == (bank.n_cosmo, self.Z, nnw)
        assert self.alias.shape == (self.S, self.Z) and self.stw.shape == (self.npar, self.S)
        self.handle = _vp()
        _check(ctx.lib.cfp_spectro_plan_create(
            ctx.handle, bank.handle, self.S, self.Z, self.Nmu, self.Nk, self.npar, _p(self.k), _p(self.mu),
            _p(self.wk), _p(self.wmu), _p(self.scal), _p(self.alias, _ip), nnw, _p(self.pnw_k), _p(self.pnw_p),
            _p(self.stw), _p(self.stden), _p(self.ps_bin, _ip), int(ps_src), _p(self.nbar), _p(self.vol),
            int(flags), int(tab_plin), int(tab_f), C.byref(self.handle)), "cfp_spectro_plan_create")
        self.h2d_bytes = sum(a.nbytes for a in (self.k, self.mu, self.wk, self.wmu, self.scal, self.alias, self.pnw_k,
                                                self.pnw_p, self.stw, self.stden, self.ps_bin, self.nbar, self.vol))

    def set_slice(self, begin: int, end: int):
        _check(self.ctx.lib.cfp_spectro_plan_set_slice(self.handle, int(begin), int(end)), "cfp_spectro_plan_set_slice")

    def run(self, materialize: int = 0, stream: int = 0):
        _check(self.ctx.lib.cfp_spectro_plan_run(self.handle, int(materialize), _vp(stream) if stream else None),
               "cfp_spectro_plan_run")

    def kernel_ms(self, part: int = 0) -> float:
        """Device time of the lattice pass of the last run: 0 whole pass, 1 weights pre-pass, 2 main kernel."""
        return float(self.ctx.lib.cfp_spectro_plan_kernel_ms_part(self.handle, int(part)))

    def fisher_device_ptr(self) -> int:
        return int(self.ctx.lib.cfp_spectro_plan_fisher_d(self.handle) or 0)

    def terms(self, s: int = 0) -> np.ndarray:
        """[5][Z][Nmu][Nk]: Kaiser term, FoG factor, de-wiggled spectrum, redshift-error factor, P_obs of stencil point s."""
        out = np.empty((5, self.Z, self.Nmu, self.Nk))
        _check(self.ctx.lib.cfp_spectro_plan_terms(self.handle, int(s), _p(out)), "cfp_spectro_plan_terms")
        return out

    def chi2(self, shot, data=None, shot_data=None, stream: int = 0, data_dev=None) -> np.ndarray:
        """Wedge chi^2 of every point of the plan against `data` (default: point 0 + noise)."""
        shot = f64(shot)
        assert shot.shape == (self.S, self.Z)
        data = None if data is None else f64(data)
        sd = None if shot_data is None else f64(shot_data)
        out = np.empty(self.S)
        if isinstance(data_dev, DeviceArray):
            assert data_dev.shape == (self.Z, self.Nmu, self.Nk) and data is None and sd is None
            _check(self.ctx.lib.cfp_spectro_plan_chi2_dev(self.handle, data_dev.ptr, _p(shot), _p(out),
                                                          _vp(stream) if stream else None), "cfp_spectro_plan_chi2_dev")
            return out
        _check(self.ctx.lib.cfp_spectro_plan_chi2(self.handle, _p(data), _p(shot), _p(sd), _p(out),
                                                  _vp(stream) if stream else None), "cfp_spectro_plan_chi2")
        return out

    def chi2_legendre(self, pl_data, icov, lw, shot, stream: int = 0) -> np.ndarray:
        """Multipole chi^2 of every point of the plan: data multipoles (Nk, Z, 3), their inverse covariance
        (Nk, Z, 3, 3), Legendre weights (3, Nmu), shot noise (S, Z)  (cfp_spectro_plan_chi2_legendre)."""
        pl_data, icov, lw, shot = f64(pl_data), f64(icov), f64(lw), f64(shot)
        assert pl_data.shape == (self.Nk, self.Z, 3) and icov.shape == (self.Nk, self.Z, 3, 3)
        assert lw.shape == (3, self.Nmu) and shot.shape == (self.S, self.Z)
        out = np.empty(self.S)
        _check(self.ctx.lib.cfp_spectro_plan_chi2_legendre(self.handle, _p(pl_data), _p(icov), _p(lw), _p(shot), _p(out),
                                                           _vp(stream) if stream else None),
               "cfp_spectro_plan_chi2_legendre")
        return out

    def fetch(self, lnp=False, derivs=False):
        F = np.empty((self.Z, self.npar, self.npar))
        L = np.empty((self.S, self.Z, self.Nmu, self.Nk)) if lnp else None
        D = np.empty((self.npar, self.Z, self.Nmu, self.Nk)) if derivs else None
        V = np.empty((self.Z, self.Nmu, self.Nk)) if derivs else None
        _check(self.ctx.lib.cfp_spectro_plan_fetch(self.handle, _p(F), _p(L), _p(D), _p(V)), "cfp_spectro_plan_fetch")
        return F, L, D, V

    def close(self):
        if self.handle:
            if self.ctx.handle:  # a destroyed context has already returned its whole arena
                self.ctx.lib.cfp_spectro_plan_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass


class PhotoPlan:
    def __init__(self, ctx: Context, bank: Bank, *, cosmo_of_s, ell, z, dz, chi, hub, wlpref, kind, ngal, bias, ia,
                 lmin, lmax, noise, sqrtfsky, stw, stden, kmax, tab_p=TAB_PNL, zmap=None, win_of_s=None, tab_pg=None,
                 ia_model=None):
        """`ia`: intrinsic-alignment window [S][Nz] of every point, or None with `ia_model` = dict(zdep, lum, growth[NC][nia],
        fac[S], eta[S], beta[S], j[Nz], t[Nz]) -- the eNLA window evaluated on the device (cfp_photo_plan_set_ia_enla)."""
        self.ctx, self.bank = ctx, bank
        self.cosmo_of_s = i32(cosmo_of_s)
        self.ell, self.z = f64(ell), f64(z)
        self.chi, self.hub, self.wlpref = f64(chi), f64(hub), f64(wlpref)
        self.kind = i32(kind)
        self.ngal, self.bias, self.ia = f64(ngal), f64(bias), (None if ia is None else f64(ia))
        self.lmin, self.lmax, self.noise, self.sqrtfsky = f64(lmin), f64(lmax), f64(noise), f64(sqrtfsky)
        self.stw, self.stden = f64(stw), f64(stden)
        self.S, self.NC = self.cosmo_of_s.size, self.chi.shape[0]
        self.Nl, self.Nz, self.Nb, self.npar = self.ell.size, self.z.size, self.kind.size, self.stw.shape[0]
        assert self.chi.shape == (self.NC, self.Nz) and self.hub.shape == (self.NC, self.Nz)
        self.zmap = None if zmap is None else i32(zmap)
        Kb = self.Nz if self.zmap is None else self.bias.shape[-1]
        self.win_of_s = None if win_of_s is None else i32(win_of_s)
        NW = 1 if self.win_of_s is None else self.ngal.shape[0]
        assert self.ngal.shape == ((self.Nb, self.Nz) if self.win_of_s is None else (NW, self.Nb, self.Nz))
        shared = self.zmap is not None and self.bias.ndim == 2  # [S][Kb]: one bias vector per point for all rows
        if shared:
            Kb = self.bias.shape[1]
            assert self.win_of_s is None and self.bias.shape == (self.S, Kb)
        else:
            assert self.bias.shape == (self.S, self.Nb, Kb)
        assert (self.ia is None or self.ia.shape == (self.S, self.Nz)) and self.stw.shape == (self.npar, self.S)
        self.handle = _vp()
        if self.win_of_s is not None:  # several sets of n(z) windows (photo-z parameters vary between points)
            assert self.win_of_s.shape == (self.S,)
            _check(ctx.lib.cfp_photo_plan_create_ex(
                ctx.handle, bank.handle, self.S, self.NC, self.Nl, self.Nz, self.Nb, self.npar,
                _p(self.cosmo_of_s, _ip), _p(self.ell), _p(self.z), float(dz), _p(self.chi), _p(self.hub),
                _p(self.wlpref), _p(self.kind, _ip), int(NW), _p(self.win_of_s, _ip), _p(self.ngal), _p(self.bias),
                int(Kb), _p(self.zmap, _ip), _p(self.ia), _p(self.lmin), _p(self.lmax), _p(self.noise),
                _p(self.sqrtfsky), _p(self.stw), _p(self.stden), float(kmax), int(tab_p), C.byref(self.handle)),
                "cfp_photo_plan_create_ex")
        elif self.zmap is None:
            _check(ctx.lib.cfp_photo_plan_create(
                ctx.handle, bank.handle, self.S, self.NC, self.Nl, self.Nz, self.Nb, self.npar,
                _p(self.cosmo_of_s, _ip), _p(self.ell), _p(self.z), float(dz), _p(self.chi), _p(self.hub),
                _p(self.wlpref), _p(self.kind, _ip), _p(self.ngal), _p(self.bias), _p(self.ia), _p(self.lmin),
                _p(self.lmax), _p(self.noise), _p(self.sqrtfsky), _p(self.stw), _p(self.stden), float(kmax), int(tab_p),
                C.byref(self.handle)), "cfp_photo_plan_create")
        else:  # bias on Kb samples per (point, row) -- or per point, shared by the rows -- + z -> sample map
            assert self.zmap.shape == (self.Nz,)
            create = ctx.lib.cfp_photo_plan_create_zmap_shared if shared else ctx.lib.cfp_photo_plan_create_zmap
            _check(create(
                ctx.handle, bank.handle, self.S, self.NC, self.Nl, self.Nz, self.Nb, self.npar,
                _p(self.cosmo_of_s, _ip), _p(self.ell), _p(self.z), float(dz), _p(self.chi), _p(self.hub),
                _p(self.wlpref), _p(self.kind, _ip), _p(self.ngal), _p(self.bias), int(Kb), _p(self.zmap, _ip),
                _p(self.ia), _p(self.lmin), _p(self.lmax), _p(self.noise), _p(self.sqrtfsky), _p(self.stw),
                _p(self.stden), float(kmax), int(tab_p), C.byref(self.handle)), "cfp_photo_plan_create_zmap")
        if tab_pg is not None and int(tab_pg) != int(tab_p):  # clustering rows on another spectrum (cdm+baryons)
            _check(ctx.lib.cfp_photo_plan_set_gc_table(self.handle, int(tab_pg)), "cfp_photo_plan_set_gc_table")
        up = [self.cosmo_of_s, self.ell, self.z, self.chi, self.hub, self.wlpref, self.kind, self.ngal, self.bias,
              self.lmin, self.lmax, self.noise, self.sqrtfsky, self.stw, self.stden]
        if self.ia is not None:
            up.append(self.ia)
        if ia_model is not None:
            m = ia_model
            zdep, lum, growth = f64(m["zdep"]).ravel(), f64(m["lum"]).ravel(), f64(m["growth"])
            fac, eta, beta = f64(m["fac"]).ravel(), f64(m["eta"]).ravel(), f64(m["beta"]).ravel()
            jj, tt = i32(m["j"]).ravel(), f64(m["t"]).ravel()
            nia = zdep.size
            assert lum.size == nia and growth.shape == (self.NC, nia) and jj.size == self.Nz and tt.size == self.Nz
            assert fac.size == self.S and eta.size == self.S and beta.size == self.S
            _check(ctx.lib.cfp_photo_plan_set_ia_enla(self.handle, nia, _p(zdep), _p(lum), _p(growth), _p(fac), _p(eta),
                                                      _p(beta), _p(jj, _ip), _p(tt)), "cfp_photo_plan_set_ia_enla")
            up += [zdep, lum, growth, fac, eta, beta, jj, tt]
        self.h2d_bytes = sum(a.nbytes for a in up)

    def set_slice(self, l_begin: int, l_end: int):
        _check(self.ctx.lib.cfp_photo_plan_set_slice(self.handle, int(l_begin), int(l_end)), "cfp_photo_plan_set_slice")

    def run(self, stream: int = 0):
        _check(self.ctx.lib.cfp_photo_plan_run(self.handle, _vp(stream) if stream else None), "cfp_photo_plan_run")

    def kernel_ms(self, which: int) -> float:
        return float(self.ctx.lib.cfp_photo_plan_kernel_ms(self.handle, int(which)))

    def fisher_device_ptr(self) -> int:
        return int(self.ctx.lib.cfp_photo_plan_fisher_d(self.handle) or 0)

    def fetch_windows(self):
        """(U, V) [S][Nb][Nz]: W_a / sqrt(H) and W^IA_a / sqrt(H) of the last run."""
        U, V = np.empty((self.S, self.Nb, self.Nz)), np.empty((self.S, self.Nb, self.Nz))
        _check(self.ctx.lib.cfp_photo_plan_fetch_windows(self.handle, _p(U), _p(V)), "cfp_photo_plan_fetch_windows")
        return U, V

    def chi2(self, blk_off, blk_n, blk_w, data=None, stream: int = 0) -> np.ndarray:
        """chi^2 of every point of the plan; blocks of tomographic rows with per-multipole weights."""
        off, n, w = i32(blk_off), i32(blk_n), f64(blk_w)
        assert w.shape == (off.size, self.Nl)
        data = None if data is None else f64(data)
        out = np.empty(self.S)
        _check(self.ctx.lib.cfp_photo_plan_chi2(self.handle, _p(data), off.size, _p(off, _ip), _p(n, _ip), _p(w), _p(out),
                                                _vp(stream) if stream else None), "cfp_photo_plan_chi2")
        return out

    def chi2_launch(self, blk_off, blk_n, blk_w, data=None, stream: int = 0) -> None:
        """Enqueue the chi^2 of every point (no waiting): prepare the next batch, then `chi2_collect()`."""
        off, n, w = i32(blk_off), i32(blk_n), f64(blk_w)
        assert w.shape == (off.size, self.Nl)
        data = None if data is None else f64(data)
        self._chi_inputs = (off, n, w, data)  # stay alive until the collect
        _check(self.ctx.lib.cfp_photo_plan_chi2_launch(self.handle, _p(data), off.size, _p(off, _ip), _p(n, _ip), _p(w),
                                                       _vp(stream) if stream else None), "cfp_photo_plan_chi2_launch")

    def chi2_collect(self) -> np.ndarray:
        out = np.empty(self.S)
        _check(self.ctx.lib.cfp_photo_plan_chi2_collect(self.handle, _p(out)), "cfp_photo_plan_chi2_collect")
        self._chi_inputs = None
        return out

    def fetch(self, cls=False, sqrtp=False):
        F = np.empty((self.npar, self.npar))
        Cl = np.empty((self.S, self.Nl, self.Nb, self.Nb)) if cls else None
        T = np.empty((self.NC, self.Nl, self.Nz)) if sqrtp else None
        _check(self.ctx.lib.cfp_photo_plan_fetch(self.handle, _p(F), _p(Cl), _p(T)), "cfp_photo_plan_fetch")
        return F, Cl, T

    def close(self):
        if self.handle:
            if self.ctx.handle:  # a destroyed context has already returned its whole arena
                self.ctx.lib.cfp_photo_plan_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):  # pragma: no cover
        try:
            if sys.is_finalizing():  # interpreter teardown: destruction order is arbitrary, the OS reclaims
                return
            self.close()
        except Exception:
            pass

"""Likelihoods on the GPU observables: same classes as the reference's `cosmicfishpie.likelihood`
(base.py `Likelihood`, photo_like.py `PhotometricLikelihood`, spectro_like.py `SpectroLikelihood` and its
module functions), plus a batched entry `loglike_batch(param_matrix, prior)` that evaluates thousands of
parameter points in one device pass (what a vectorised Nautilus `Sampler(..., vectorized=True)` calls).

A `FisherMatrix` instance is used purely as configuration + fiducial-data carrier, as in the reference.
The shipped reference `SpectroLikelihood` cannot be constructed (it reads `self.cosmoFM_data`, which its base
class never sets, and `pk_cov.volume_survey_array`, which does not exist; SURVEY.md §8a-L2'); here both work.

Cosmological parameters of a point are mapped to the nearest tabulated cosmology of the table bank, exactly
like the reference's `external` input does (cosmology.py:1375-1421); nuisance parameters are continuous.
"""
from __future__ import annotations

import os

from abc import ABC, abstractmethod
from copy import deepcopy
from typing import Any, Dict, Iterable, Optional, Sequence

import numpy as np

from . import _lib
from . import photo as photo_mod
from . import spectro as spectro_mod
from .quadrature import simpson_weights


def _dict_with_updates(template: Dict[str, Any], pool: Dict[str, Any]) -> Dict[str, Any]:
    out = deepcopy(template)
    for key in template:
        if key in pool:
            out[key] = pool.pop(key)
    return out


class Likelihood(ABC):
    """Reference: likelihood/base.py:33-173."""

    def __init__(self, *, cosmo_data, cosmo_theory=None, leg_flag: str = "wedges") -> None:
        self.cosmo_data = cosmo_data
        self.cosmo_theory = cosmo_theory or cosmo_data
        self.cosmoFM_data, self.cosmoFM_theory = self.cosmo_data, self.cosmo_theory
        self.leg_flag = leg_flag
        self.data_obs = self.compute_data()

    @abstractmethod
    def compute_data(self) -> Any: ...

    @abstractmethod
    def compute_theory(self, param_dict: Dict[str, Any]) -> Any: ...

    @abstractmethod
    def compute_chi2(self, theory_obs: Any) -> float: ...

    @abstractmethod
    def chi2_batch(self, param_dicts: Sequence[Dict[str, Any]]) -> np.ndarray: ...

    def build_param_dict(self, *, param_vec=None, param_dict=None, prior=None) -> Dict[str, Any]:
        if param_dict is not None:
            return dict(param_dict)
        if param_vec is None or prior is None:
            raise ValueError("Provide either param_dict or (param_vec and prior) to build the parameter mapping")
        if not (isinstance(param_vec, (list, tuple, np.ndarray)) and not isinstance(param_vec, (str, bytes))):
            raise TypeError("param_vec must be an indexable iterable when no param_dict is supplied")
        keys = getattr(prior, "keys", None)
        if keys is None:
            raise AttributeError("prior object must expose an ordered 'keys' attribute to map the vector to parameters")
        return {key: param_vec[i] for i, key in enumerate(keys)}

    def loglike(self, param_vec=None, *, param_dict=None, prior=None) -> float:
        params = self.build_param_dict(param_vec=param_vec, param_dict=param_dict, prior=prior)
        return float(-0.5 * self.chi2_batch([params])[0])

    def chi2_matrix(self, names, matrix) -> np.ndarray:
        """chi^2 of every row of `matrix` [B, len(names)]; subclasses override this with an array fast path."""
        return self.chi2_batch([dict(zip(names, row)) for row in matrix])

    def loglike_batch(self, param_matrix, prior=None, keys: Optional[Iterable[str]] = None, shard: bool = False) -> np.ndarray:
        """log-likelihood of every row of `param_matrix` [B, ndim]; columns named by `prior.keys` or `keys`.

        `shard=True` under torch.distributed (one process per GPU): every rank holds the same matrix, evaluates a
        contiguous slice of the points on its own GPU and ONE all-gather (NCCL over NVLink; gloo on CPU) hands every
        rank the full vector -- the points are independent, so this is the only exchange."""
        names = list(keys) if keys is not None else list(getattr(prior, "keys"))
        pm = np.atleast_2d(np.asarray(param_matrix, dtype=float))
        if shard:
            from . import distributed as cdist

            rank, world = cdist.rank_world()
            if world > 1:
                B = pm.shape[0]
                per = (B + world - 1) // world  # equal slices (the last ranks may run past the end: padded below)
                lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
                local = np.full(per, np.nan)
                if hi > lo:
                    local[: hi - lo] = -0.5 * self.chi2_matrix(names, pm[lo:hi])
                dev = (getattr(self.cosmo_theory, "settings", None) or {}).get("device")
                return cdist.allgather_numpy(local, device=dev)[:B]
        return -0.5 * self.chi2_matrix(names, pm)


class NautilusMixin:
    """Reference: likelihood/base.py:176-220 (nautilus is imported lazily; it is an optional dependency)."""

    def create_nautilus_prior(self, prior_dict):
        from nautilus import Prior

        prior = Prior()
        for par, (lo, hi) in prior_dict.items():
            prior.add_parameter(par, (lo, hi))
        return prior

    def run_nautilus(self, *, prior, sampler_kwargs=None, run_kwargs=None):
        from nautilus import Sampler

        kw = dict(sampler_kwargs or {})
        kw.pop("pool", None)  # never fork after CUDA initialisation: batches go to the GPU instead
        sampler = Sampler(prior, lambda x: self.loglike_batch(x, prior=prior, shard=True), vectorized=True, pass_dict=False, **kw)
        sampler.run(**dict(run_kwargs or {}))
        return sampler


# ------------------------------------------------------------------------------------- photometric
def quadrature_weights(ells, limit):
    """w[l] such that sum_l w[l] q[l] is the reference's sum over multipoles of one chi^2 block (photo_like.py:90-96
    with the cut of :189-204): the multipoles below `limit` weighted by (2 l + 1), integrated with the trapezoid rule
    on their own spacings plus one closing rectangle whose width runs from the last kept multipole to the cut itself
    (without a cut: the previous spacing once more)."""
    ells = np.asarray(ells, dtype=float)
    w = np.zeros(ells.size)
    n = ells.size if limit is None else int(np.searchsorted(ells, limit))
    if n == 0:
        return w
    steps = np.diff(ells[:n])
    if limit is not None:
        closing = limit - ells[n - 1]
    else:
        closing = steps[-1] if steps.size else 1.0
    w[: n - 1] += steps / 2
    w[1:n] += steps / 2
    w[n - 1] += closing
    w[:n] *= 2 * ells[:n] + 1
    return w


class PhotometricLikelihood(Likelihood, NautilusMixin):
    """Reference: likelihood/photo_like.py:100-256."""

    def __init__(self, *, cosmo_data, cosmo_theory=None, observables=None, data_cells=None) -> None:
        self.observables = tuple(observables or cosmo_data.observables)
        self._preloaded_cells = None if data_cells is None else {k: np.array(v) for k, v in data_cells.items()}
        self.photo_cov_data = None
        super().__init__(cosmo_data=cosmo_data, cosmo_theory=cosmo_theory, leg_flag="cells")

    # --- helpers ---------------------------------------------------------------------------------
    def _grid(self):
        return photo_mod._photo_grid(self.cosmo_data)

    def _cells_from_full(self, cl_noisy):
        """[Nl][Nb][Nb] (noise included) -> {"ells", Cell_LL, Cell_GG, Cell_GL} (photo_like.py:28-75)."""
        g = self._g
        out = {"ells": np.array(g.ell, copy=True)}
        wl = [a for a, (o, _) in enumerate(g.rows) if o == "WL"]
        gc = [a for a, (o, _) in enumerate(g.rows) if o == "GCph"]
        if wl and "WL" in self.observables:
            out["Cell_LL"] = cl_noisy[:, wl][:, :, wl]
        if gc and "GCph" in self.observables:
            out["Cell_GG"] = cl_noisy[:, gc][:, :, gc]
        if wl and gc and "WL" in self.observables and "GCph" in self.observables:
            out["Cell_GL"] = cl_noisy[:, gc][:, :, wl]
        return out

    def _full_from_cells(self, cells):
        g = self._g
        n = len(g.rows)
        full = np.zeros((len(g.ell), n, n))
        wl = [a for a, (o, _) in enumerate(g.rows) if o == "WL"]
        gc = [a for a, (o, _) in enumerate(g.rows) if o == "GCph"]
        if "Cell_LL" in cells:
            full[np.ix_(range(len(g.ell)), wl, wl)] = cells["Cell_LL"]
        if "Cell_GG" in cells:
            full[np.ix_(range(len(g.ell)), gc, gc)] = cells["Cell_GG"]
        if "Cell_GL" in cells:
            full[np.ix_(range(len(g.ell)), gc, wl)] = cells["Cell_GL"]
            full[np.ix_(range(len(g.ell)), wl, gc)] = np.transpose(cells["Cell_GL"], (0, 2, 1))
        return full

    def _blocks(self):
        """Row blocks and per-multipole weights of photo_like.py:206-254 (the packaged composition, incl.
        the second LL term on the cross-correlation range)."""
        g, sp = self._g, self.cosmo_data.specs
        fw, fg = sp.get("fsky_WL", 1.0), sp.get("fsky_GCph", 1.0)
        lw, lg = sp.get("lmax_WL"), sp.get("lmax_GCph")
        lx = min(lw, lg) if (lw is not None and lg is not None) else None
        wl = [a for a, (o, _) in enumerate(g.rows) if o == "WL"]
        gc = [a for a, (o, _) in enumerate(g.rows) if o == "GCph"]
        off, n, w = [], [], []
        has_wl, has_gc = bool(wl) and "WL" in self.observables, bool(gc) and "GCph" in self.observables
        if has_wl:
            wt = fw * quadrature_weights(g.ell, lw)
            if has_gc:
                wt = wt + fw * quadrature_weights(g.ell, lx)
            off.append(wl[0]); n.append(len(wl)); w.append(wt)
        if has_gc:
            off.append(gc[0]); n.append(len(gc)); w.append(fg * quadrature_weights(g.ell, lg))
        if has_wl and has_gc:
            off.append(0); n.append(len(g.rows)); w.append(np.sqrt(fw * fg) * quadrature_weights(g.ell, lx))
        return np.array(off, np.int32), np.array(n, np.int32), np.array(w)

    # --- reference interface -----------------------------------------------------------------------
    def compute_data(self):
        self._g = self._grid()
        fm = self.cosmo_data
        self._ellmax_WL, self._ellmax_GC = fm.specs.get("lmax_WL"), fm.specs.get("lmax_GCph")
        if self._preloaded_cells is not None:
            self._ells = np.array(self._preloaded_cells.get("ells"), copy=True)
            self._data_full = self._full_from_cells(self._preloaded_cells)
            return self._preloaded_cells
        self.photo_cov_data = getattr(fm, "photo_LSS", None)
        if self.photo_cov_data is None:
            self.photo_cov_data = photo_mod.PhotoCov(fm.fiducialcosmopars, fm.photopars, fm.IApars, fm.photobiaspars,
                                                     configuration=fm)
        fid = {**fm.fiducialcosmopars, **fm.IApars, **fm.photobiaspars, **fm.photopars}
        job = photo_mod.PhotoJob(fm, [fid], np.zeros((1, 1)), np.ones(1))
        _, cl, _ = job.run().fetch(cls=True)
        job.close()
        noisy = cl[0] + np.eye(len(self._g.rows))[None] * self._g.noise[None, None, :]
        self._data_full = noisy
        cells = self._cells_from_full(noisy)
        self._ells = cells["ells"]
        return cells

    def _allpars(self, param_dict):
        fm = self.cosmo_theory
        p = deepcopy(param_dict)
        cp = _dict_with_updates(fm.fiducialcosmopars, p)
        pp = _dict_with_updates(fm.photopars, p)
        ia = _dict_with_updates(fm.IApars, p)
        bp = _dict_with_updates(fm.photobiaspars, p)
        return {**cp, **ia, **bp, **pp}

    def compute_theory(self, param_dict):
        job = photo_mod.PhotoJob(self.cosmo_theory, [self._allpars(param_dict)], np.zeros((1, 1)), np.ones(1))
        _, cl, _ = job.run().fetch(cls=True)
        job.close()
        return self._cells_from_full(cl[0] + np.eye(len(self._g.rows))[None] * self._g.noise[None, None, :])

    def compute_chi2(self, theory_obs) -> float:
        off, n, w = self._blocks()
        ctx = _lib.default_context(self.cosmo_data.settings.get("device"))
        return float(ctx.cells_chi2(self._full_from_cells(theory_obs)[None], self._data_full, off, n, w)[0])

    def chi2_matrix(self, names, matrix) -> np.ndarray:
        fmt = self.cosmo_theory
        known = set(fmt.fiducialcosmopars) | set(fmt.IApars) | set(fmt.photobiaspars) | set(fmt.photopars)
        keep = [i for i, n in enumerate(names) if n in known]  # unused parameters are ignored, as in the reference
        names = [names[i] for i in keep]
        matrix = np.ascontiguousarray(np.atleast_2d(matrix)[:, keep])
        if any(n in fmt.photopars for n in names):
            # photo-z parameters change the n_i(z) windows of a point (photo_like.py compute_theory updates photopars):
            # the array fast path has one window set per job, so these batches go point by point
            return self.chi2_batch([dict(zip(names, row)) for row in matrix])
        off, n, w = self._blocks()
        out = np.empty(matrix.shape[0])
        chunk = 4096
        spans = [(i, min(i + chunk, matrix.shape[0])) for i in range(0, matrix.shape[0], chunk)]
        # (between batches, never between the pipelined chunks of one: a running chunk holds the bank)
        photo_mod._cosmo_set_of(photo_mod._cfg_of(fmt)[1]).begin_batch()
        if len(spans) == 1:
            job = photo_mod.PhotoJob.from_matrix(self.cosmo_theory, names, matrix)
            out[:] = job.plan().chi2(off, n, w, data=self._data_full)
            job.close()
            return out
        # software pipeline over the chunks, no threads: chunk i's kernels are ENQUEUED on one of the context's two extra
        # streams in turn (cfp_photo_plan_chi2_launch returns without waiting) and run while this thread assembles and
        # uploads the next chunks; two chunks are kept in flight, so the C_ell contraction of one (tensor path) overlaps
        # the factorisations of the other (bound by instruction issue); results are collected two chunks later
        from collections import deque

        pending = deque()
        depth = int(os.environ.get("CFP_LIKE_PIPELINE_DEPTH", "2"))

        def collect_oldest():
            pjob, pa, pb = pending.popleft()
            out[pa:pb] = pjob.plan().chi2_collect()
            pjob.close()

        try:
            for i, (a, b) in enumerate(spans):
                job = photo_mod.PhotoJob.from_matrix(self.cosmo_theory, names, matrix[a:b])
                plan = job.plan()  # create + H2D, on the context stream
                plan.chi2_launch(off, n, w, data=self._data_full,
                                 stream=plan.ctx.aux_stream if i % 2 == 0 else plan.ctx.aux_stream2)
                pending.append((job, a, b))
                while len(pending) > depth:
                    collect_oldest()
            while pending:
                collect_oldest()
        finally:
            for pjob, _, _ in pending:  # (only after an error: chunks still in flight give their buffers back)
                try:
                    pjob.close()
                except Exception:  # noqa: BLE001
                    pass
        return out

    def chi2_batch(self, param_dicts) -> np.ndarray:
        points = [self._allpars(p) for p in param_dicts]
        off, n, w = self._blocks()
        out = np.empty(len(points))
        chunk = 4096
        for i in range(0, len(points), chunk):
            job = photo_mod.PhotoJob(self.cosmo_theory, points[i:i + chunk], np.zeros((1, len(points[i:i + chunk]))),
                                     np.ones(1))
            out[i:i + chunk] = job.plan().chi2(off, n, w, data=self._data_full)
            job.close()
        return out


# ------------------------------------------------------------------------------------- spectroscopic
def _device_of(fm):
    """options["device"] of a FisherMatrix; duck-typed stand-ins (the reference's own tests use bare objects with
    only the grids) have no settings and get the default device."""
    return (getattr(fm, "settings", None) or {}).get("device") if fm is not None else None


def observable_Pgg(theory_spectro, cosmoFM, nuisance_shot=None):
    """P_obs + 1/nbar + shot on the (z, mu, k) lattice (reference: spectro_like.py:25-41)."""
    z_bins = cosmoFM.pk_cov.global_z_bin_mids
    if nuisance_shot is None:
        nuisance_shot = np.zeros_like(z_bins)
    pobs = theory_spectro.pk_obs
    lat = np.exp(spectro_mod._evaluate_lattice([pobs], list(z_bins), cosmoFM.Pk_kgrid, cosmoFM.Pk_mugrid))[0]
    for i, z in enumerate(z_bins):
        lat[i] += 1 / theory_spectro.n_density(z) + nuisance_shot[i]
    return lat


def compute_wedge_chi2(P_obs_data, P_obs_theory, cosmoFM_data):
    """Reference: spectro_like.py:143-189 (device reduction, Simpson weights of the lattice)."""
    ctx = _lib.default_context(_device_of(cosmoFM_data))
    k, mu = cosmoFM_data.Pk_kgrid, cosmoFM_data.Pk_mugrid
    return float(ctx.wedge_chi2(np.asarray(P_obs_theory)[None], P_obs_data, k, simpson_weights(k), simpson_weights(mu),
                                cosmoFM_data.pk_cov.volume_survey_array)[0])


# Wigner-3j sum matrices for the multipole covariance (utilities/legendre_tools.py:56-123): m00, m22, m44, m02, m24, m04
_M3J_REF = np.array([  # with the reference's 15-digit literals, so results agree to the last bits
    [1.00000000000000, 0, 0, 0, 0.200000000000000, 0, 0, 0, 0.111111111111111],
    [0.200000000000000, 2 / 35, 2 / 35, 2 / 35, 0.0857142857142857, 12 / 385, 2 / 35, 12 / 385, 0.0397158397158397],
    [0.111111111111111, 20 / 693, 18 / 1001, 20 / 693, 0.0397158397158397, 20 / 1001, 18 / 1001, 20 / 1001, 0.0310865604983252],
    [0, 0.200000000000000, 0, 0.200000000000000, 2 / 35, 0.0571428571428571, 0, 0.0571428571428571, 20 / 693],
    [0, 0.0571428571428571, 20 / 693, 0.0571428571428571, 12 / 385, 0.0397158397158397, 20 / 693, 0.0397158397158397, 20 / 1001],
    [0, 0, 0.111111111111111, 0, 2 / 35, 20 / 693, 0.111111111111111, 20 / 693, 18 / 1001],
])
leg_ells = np.arange(0, 5, 2, dtype="int")


def _legendre_weights(mugrid):
    """(2 ell + 1) * Simpson weight * L_ell(mu) for ell = 0, 2, 4 (spectro_like.py:49-54)."""
    from scipy.special import eval_legendre

    w = simpson_weights(mugrid)
    return np.array([(2.0 * ell + 1) * w * eval_legendre(ell, mugrid) for ell in leg_ells])


def legendre_Pgg(obs_Pgg, cosmoFM):
    """Multipoles P_ell(k, z), ell = 0,2,4, shape (Nk, Nz, 3) (reference: spectro_like.py:48-56)."""
    ctx = _lib.default_context(_device_of(cosmoFM))
    return ctx.legendre_project(np.asarray(obs_Pgg)[None], _legendre_weights(cosmoFM.Pk_mugrid))[0]


def compute_covariance_legendre(P_ell, cosmoFM):
    """k-bin integrated Gaussian covariance of the multipoles and its inverse (reference: spectro_like.py:69-122)."""
    ctx = _lib.default_context(_device_of(cosmoFM))
    if P_ell.shape[0] != len(cosmoFM.Pk_kgrid):
        raise ValueError("k_grid and P_ell have different lengths in k")
    return ctx.legendre_covariance(P_ell, cosmoFM.Pk_kgrid, cosmoFM.pk_cov.volume_survey_array, _M3J_REF)


def compute_chi2_legendre(P_ell_data, P_ell_theory, inv_covariance, cosmoFM=None):
    """Reference: spectro_like.py:125-140."""
    ctx = _lib.default_context(_device_of(cosmoFM))
    return float(ctx.legendre_chi2(P_ell_data, np.asarray(P_ell_theory)[None], inv_covariance)[0])


def _spectro_point(param_dict, fm):
    """(ComputeGalSpectro of the point, shot noise per bin)  (reference: spectro_like.py:192-224)."""
    p = deepcopy(param_dict)
    z_bins = fm.pk_cov.global_z_bin_mids
    shot = np.zeros(len(z_bins))
    for i, key in enumerate(fm.PShotpars.keys()):
        shot[i] = p.pop(key, fm.PShotpars[key])
    bp = _dict_with_updates(fm.Spectrobiaspars, p)
    nl = _dict_with_updates(fm.Spectrononlinpars, p)
    cp = _dict_with_updates(fm.fiducialcosmopars, p)
    obs = spectro_mod.ComputeGalSpectro(cp, spectrobiaspars=bp, spectrononlinearpars=nl, PShotpars=fm.PShotpars,
                                        configuration=fm, cosmo_set=fm.pk_obs_fid.cosmo_set)
    return obs, shot


def compute_theory_spectro(param_dict, cosmoFM_theory, leg_flag="wedges"):
    obs, shot = _spectro_point(param_dict, cosmoFM_theory)
    cov = spectro_mod.SpectroCov(obs.cosmopars, configuration=cosmoFM_theory, fiducial_specobs=obs)
    out = observable_Pgg(cov, cosmoFM_theory, nuisance_shot=shot)
    if leg_flag == "wedges":
        return out
    if leg_flag == "legendre":
        return legendre_Pgg(out, cosmoFM_theory)
    raise ValueError(f"Unknown leg_flag '{leg_flag}'. Use 'wedges' or 'legendre'.")


class SpectroLikelihood(Likelihood, NautilusMixin):
    """Reference: likelihood/spectro_like.py:236-299 (wedges and Legendre multipoles)."""

    def __init__(self, *, cosmoFM_data, cosmoFM_theory=None, leg_flag="wedges", data_obs=None, nuisance_shot=None):
        if leg_flag not in ("wedges", "legendre"):
            raise ValueError(f"Unknown leg_flag '{leg_flag}'. Use 'wedges' or 'legendre'.")
        self._inv_cov_legendre = None
        self._data_wedges = None
        self._preloaded_data = None if data_obs is None else np.array(data_obs)
        self._nuisance_shot = None if nuisance_shot is None else np.array(nuisance_shot, dtype=float)
        super().__init__(cosmo_data=cosmoFM_data, cosmo_theory=cosmoFM_theory, leg_flag=leg_flag)

    @property
    def data_wedges(self):
        return self._data_wedges

    def _data_on_device(self, ctx):
        """The wedge data lattice, uploaded once per device context and kept there for every batch."""
        hit = self.__dict__.get("_data_dev")
        if hit is None or hit[0] is not ctx or hit[1] is not self._data_wedges:
            hit = (ctx, self._data_wedges, _lib.DeviceArray(ctx, self._data_wedges))
            self.__dict__["_data_dev"] = hit
        return hit[2]

    def compute_data(self):
        fm = self.cosmo_data
        if not hasattr(fm, "pk_cov") or fm.pk_cov is None:
            raise AttributeError("cosmoFM_data.pk_cov is not available. Ensure the FisherMatrix was initialised for "
                                 "spectroscopic probes and compute() was called.")
        if self._preloaded_data is not None:
            data = self._preloaded_data
            if self.leg_flag == "legendre":
                _, self._inv_cov_legendre = compute_covariance_legendre(data, fm)
            else:
                self._data_wedges = data
            return data
        obs = observable_Pgg(fm.pk_cov, fm, nuisance_shot=self._nuisance_shot)
        self._data_wedges = obs
        if self.leg_flag == "legendre":
            p_ell = legendre_Pgg(obs, fm)
            _, self._inv_cov_legendre = compute_covariance_legendre(p_ell, fm)
            return p_ell
        return obs

    def compute_theory(self, param_dict):
        return compute_theory_spectro(param_dict, self.cosmo_theory, self.leg_flag)

    def compute_chi2(self, theory_obs) -> float:
        if self.leg_flag == "wedges":
            return compute_wedge_chi2(self.data_obs, theory_obs, self.cosmo_data)
        return compute_chi2_legendre(self.data_obs, theory_obs, self._inv_cov_legendre, self.cosmo_data)

    def _chi2_batch_legendre(self, param_dicts) -> np.ndarray:
        """Multipole chi^2: theory lattices of a chunk of points on the device, projected and contracted there."""
        fm = self.cosmo_theory
        zs = list(fm.pk_cov.global_z_bin_mids)
        ctx = _lib.default_context(fm.settings.get("device"))
        lw = _legendre_weights(fm.Pk_mugrid)
        inv_n = np.array([1 / fm.pk_cov.n_density(z) for z in zs])
        out = np.empty(len(param_dicts))
        if fm.Pk_musamp <= 256:  # fused: the theory lattices are never materialised
            for i0 in range(0, len(param_dicts), 4096):
                pts, shots = zip(*[_spectro_point(p, fm) for p in param_dicts[i0:i0 + 4096]])
                cs = pts[0].cosmo_set
                S, Z = len(pts), len(zs)
                spectro_mod.prepare_cosmologies(cs, zs, pts[0])
                scal = np.array([[p.scalar_block(z) for z in zs] for p in pts])
                used = {int(c) for c in scal[:, :, _lib.SP_COSMO].ravel()}
                pk, pp = spectro_mod._pnw_arrays(cs, zs, set() if pts[0].linear_switch else used)
                tabs = spectro_mod.table_slots(fm.settings)
                plan = _lib.SpectroPlan(
                    cs.context(), cs.bank(), k=fm.Pk_kgrid, mu=fm.Pk_mugrid,
                    wk=simpson_weights(fm.Pk_kgrid), wmu=simpson_weights(fm.Pk_mugrid), scal=scal,
                    alias=np.tile(np.arange(S, dtype=np.int32)[:, None], (1, Z)), pnw_k=pk, pnw_p=pp,
                    stw=np.zeros((1, S)), stden=np.ones(1), ps_bin=np.array([-1], np.int32), ps_src=0,
                    nbar=1.0 / inv_n, vol=fm.pk_cov.volume_survey_array, flags=pts[0].flags(), tab_plin=tabs[0],
                    tab_f=tabs[1])
                out[i0:i0 + S] = plan.chi2_legendre(self.data_obs, self._inv_cov_legendre, lw, np.array(shots))
                plan.close()
            return out
        lattice_bytes = 8 * len(zs) * fm.Pk_musamp * fm.Pk_ksamp
        chunk = max(1, min(512, int(2**30 // lattice_bytes)))
        for i0 in range(0, len(param_dicts), chunk):
            pts, shots = zip(*[_spectro_point(p, fm) for p in param_dicts[i0:i0 + chunk]])
            lat = np.exp(spectro_mod._evaluate_lattice(list(pts), zs, fm.Pk_kgrid, fm.Pk_mugrid))
            lat += (inv_n[None, :] + np.array(shots))[:, :, None, None]
            out[i0:i0 + chunk] = ctx.legendre_chi2(self.data_obs, ctx.legendre_project(lat, lw), self._inv_cov_legendre)
        return out

    def chi2_matrix(self, names, matrix) -> np.ndarray:
        """Array fast path of the wedge chi^2 for a batch given as a matrix: per-point scalar blocks are assembled
        from one cached block per (cosmology, bin) plus the nuisance columns, no Python object per point."""
        fm = self.cosmo_theory
        if self.leg_flag != "wedges" and fm.Pk_musamp > 256:  # the fused multipole kernel stages <= 256 mu rows
            return self.chi2_batch([dict(zip(names, row)) for row in matrix])
        fid_obs = fm.pk_obs_fid
        cs = fid_obs.cosmo_set
        cs.begin_batch()
        zs = list(fm.pk_cov.global_z_bin_mids)
        matrix = np.atleast_2d(np.asarray(matrix, dtype=float))
        B, Z = matrix.shape[0], len(zs)
        if B == 0 or matrix.size == 0:  # an empty batch (a sampler may propose none): nothing to evaluate
            return np.empty(0)
        scal, shot, cosmo_of_s = spectro_mod.scalar_blocks_from_matrix(fm, fid_obs, zs, names, matrix)
        used = {int(c) for c in np.unique(cosmo_of_s)}
        pk, pp = spectro_mod._pnw_arrays(cs, zs, set() if fid_obs.linear_switch else used)
        wk, wmu = simpson_weights(fm.Pk_kgrid), simpson_weights(fm.Pk_mugrid)
        nbar = np.array([fm.pk_cov.n_density(z) for z in zs])
        out = np.empty(B)
        chunk = 8192  # (16384- and 32768-point plans were measured slower: the upload of the scalar blocks grows faster)
        for i0 in range(0, B, chunk):
            S = min(chunk, B - i0)
            plan = _lib.SpectroPlan(
                cs.context(), cs.bank(), k=fm.Pk_kgrid, mu=fm.Pk_mugrid, wk=wk, wmu=wmu,
                scal=scal[i0:i0 + S], alias=np.tile(np.arange(S, dtype=np.int32)[:, None], (1, Z)), pnw_k=pk, pnw_p=pp,
                stw=np.zeros((1, S)), stden=np.ones(1), ps_bin=np.array([-1], np.int32), ps_src=0, nbar=nbar,
                vol=fm.pk_cov.volume_survey_array, flags=fid_obs.flags(),
                tab_plin=spectro_mod.table_slots(fm.settings)[0], tab_f=spectro_mod.table_slots(fm.settings)[1])
            if self.leg_flag == "wedges":
                out[i0:i0 + S] = plan.chi2(shot[i0:i0 + S], data_dev=self._data_on_device(plan.ctx))
            else:
                out[i0:i0 + S] = plan.chi2_legendre(self.data_obs, self._inv_cov_legendre, _legendre_weights(fm.Pk_mugrid),
                                                    shot[i0:i0 + S])
            plan.close()
        return out

    def chi2_batch(self, param_dicts) -> np.ndarray:
        if self.leg_flag == "legendre":
            return self._chi2_batch_legendre(param_dicts)
        fm = self.cosmo_theory
        zs = list(fm.pk_cov.global_z_bin_mids)
        out = np.empty(len(param_dicts))
        chunk = 8192
        for i0 in range(0, len(param_dicts), chunk):
            sub = param_dicts[i0:i0 + chunk]
            pts, shots = zip(*[_spectro_point(p, fm) for p in sub])
            cs = pts[0].cosmo_set
            S, Z = len(pts), len(zs)
            spectro_mod.prepare_cosmologies(cs, zs, pts[0])
            scal = np.array([[p.scalar_block(z) for z in zs] for p in pts])
            used = {int(c) for c in scal[:, :, _lib.SP_COSMO].ravel()}
            pk, pp = spectro_mod._pnw_arrays(cs, zs, set() if pts[0].linear_switch else used)
            plan = _lib.SpectroPlan(
                cs.context(), cs.bank(), k=fm.Pk_kgrid, mu=fm.Pk_mugrid,
                wk=simpson_weights(fm.Pk_kgrid), wmu=simpson_weights(fm.Pk_mugrid), scal=scal,
                alias=np.tile(np.arange(S, dtype=np.int32)[:, None], (1, Z)), pnw_k=pk, pnw_p=pp, stw=np.zeros((1, S)),
                stden=np.ones(1), ps_bin=np.array([-1], np.int32), ps_src=0,
                nbar=np.array([fm.pk_cov.n_density(z) for z in zs]), vol=fm.pk_cov.volume_survey_array,
                flags=pts[0].flags(), tab_plin=spectro_mod.table_slots(fm.settings)[0],
                tab_f=spectro_mod.table_slots(fm.settings)[1])
            out[i0:i0 + chunk] = plan.chi2(np.array(shots), data_dev=self._data_on_device(plan.ctx))
            plan.close()
        return out


def loglike(param_vec=None, theory_obsPgg=None, prior=None, leg_flag="wedges", cosmoFM_theory=None, cosmoFM_data=None,
            data_obsPgg=None):
    """Module-level helper of the reference (spectro_like.py:302-332)."""
    if cosmoFM_data is None:
        return -np.inf
    like = SpectroLikelihood(cosmoFM_data=cosmoFM_data, cosmoFM_theory=cosmoFM_theory, leg_flag=leg_flag,
                             data_obs=data_obsPgg)
    if theory_obsPgg is not None:
        return -0.5 * like.compute_chi2(theory_obsPgg)
    try:
        params = dict(param_vec) if isinstance(param_vec, dict) else like.build_param_dict(param_vec=param_vec, prior=prior)
    except (TypeError, ValueError, AttributeError):
        return -np.inf
    return like.loglike(param_dict=params)

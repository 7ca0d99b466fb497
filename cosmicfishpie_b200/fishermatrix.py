"""`FisherMatrix(...).compute()` -- the public entry point, same signature as the reference
(cosmicfishpie/fishermatrix/cosmicfish.py:43-364) for the observables GCsp, WL, GCph and WL+GCph.

The class keeps the attributes other code reads off a reference instance:
  photo   : photo_obs_fid, photo_LSS, photo_derivs
  spectro : pk_obs_fid, pk_cov, pk_deriv, derivs_dict, veff_arr, zmids, obs_spectrum,
            Pk_kgrid/Pk_mugrid/Pk_kmesh/Pk_mumesh/Pk_ksamp/Pk_musamp, kmin_fish/kmax_fish
  both    : settings, specs, observables, freeparams, allparams, fiducialcosmopars, fiducialcosmo, ...
`derivs_dict` and `veff_arr` (hundreds of MB on fine lattices) are materialised lazily: compute() runs the
fused kernel, which never writes the per-parameter lattices to HBM; touching either attribute re-runs the
kernel in materialising mode.

Extra options understood here (all optional, defaults reproduce the reference):
  options["stencil"]   "3PT" (default) | "5PT" | "4PT_FWD" | "STEM" (photometric probes only)  -- the reference
                       ignores settings["derivatives"] and always uses 3PT, so a separate key is used
  options["device"]    CUDA device index (default LOCAL_RANK or 0)
"""
from __future__ import annotations

import copy
import json
import pickle
import os
import subprocess
import sys
from time import time

import numpy as np

from . import _lib
from . import config as config_module
from . import distributed as cdist
from . import fisher_matrix as fm
from . import spectro as spectro_mod

VERSION = "1.3.0"  # the reference release whose results this package reproduces


def _write_now(path, text):
    """One open / write / close on the file descriptor (no buffered text layer: four small files per Fisher matrix)."""
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o666)
    try:
        data = text.encode()
        while data:
            data = data[os.write(fd, data):]
    finally:
        os.close(fd)


class _DeferredWrites:
    """Result files of a batch of forecasts (`FisherMatrix.compute_many`) written by ONE helper thread: the eight
    open / write / close system calls per forecast release the interpreter lock, so they overlap the assembly of the
    next forecast.  `flush()` returns when every file handed over so far is on disk and re-raises the first error."""

    def __init__(self):
        import queue
        import threading

        self.q, self.err = queue.Queue(), None
        self.thread = threading.Thread(target=self._run, name="cfp-export", daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.q.get()
            try:
                if item is None:
                    return
                if self.err is None:
                    _write_now(*item)
            except Exception as exc:  # noqa: BLE001 -- reported by flush()
                self.err = exc
            finally:
                self.q.task_done()

    def put(self, path, text):
        self.q.put((path, text))

    def flush(self):
        self.q.join()
        if self.err is not None:
            err, self.err = self.err, None
            raise err

    def close(self):
        self.q.put(None)
        self.thread.join()


_deferred = None  # set by compute_many for the duration of a batch


def _write_text(path, text):
    if _deferred is not None:
        _deferred.put(path, text)
    else:
        _write_now(path, text)


_ARRAY_STR = {}


def _array_str(a):
    """str(a) of a specification array, remembered by content: the same few bin-edge arrays are printed into every
    `_specs.dat`."""
    key = (a.shape, a.dtype.str, a.tobytes())
    hit = _ARRAY_STR.get(key)
    if hit is None:
        if len(_ARRAY_STR) > 256:
            _ARRAY_STR.clear()
        hit = _ARRAY_STR[key] = str(a)
    return hit


def _json_default(o):
    """What the C encoder of `json` cannot serialise by itself."""
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.floating, np.integer, np.bool_)):
        return o.item()
    if isinstance(o, range):
        return list(o)
    raise TypeError(type(o).__name__)


def _jsonable(o):
    """`o` as the encoder can take it: in-memory table banks are not written, keys become strings.  The common case --
    dictionaries whose keys are strings all the way down -- passes through untouched (the C encoder walks it)."""
    def plain(x):
        if isinstance(x, dict):
            # (the encoder itself writes an int key as str(key))
            return all((type(k) is str or type(k) is int) and k != "tables" and plain(v) for k, v in x.items())
        if isinstance(x, (list, tuple)):
            return all(plain(v) for v in x)
        return True

    def convert(x):
        if isinstance(x, dict):
            return {str(k): convert(v) for k, v in x.items() if k != "tables"}
        if isinstance(x, (list, tuple)):
            return [convert(v) for v in x]
        return x

    return o if plain(o) else convert(o)


class FisherMatrix:
    cf_version = f"CosmicFish_v{VERSION}"

    def __init__(self, options=dict(), specifications=dict(), observables=None, freepars=None, extfiles=None,
                 fiducialpars=None, photobiaspars=None, photopars=None, spectrononlinearpars=None,
                 spectrobiaspars=None, IApars=None, IMbiaspars=None, PShotpars=None, surveyName="Euclid",
                 cosmoModel="w0waCDM", latexnames=None, parallel=False, parallel_cpus=4):
        if options["feedback"] > 0:  # KeyError without "feedback", like the reference (cosmicfish.py:130)
            print("**** cosmicfishpie_b200: B200-native observable -> Fisher hot path ****")
            sys.stdout.flush()
        cfg = config_module.Config(
            options=options, specifications=specifications, observables=observables, freepars=freepars,
            extfiles=extfiles, fiducialpars=fiducialpars, photobiaspars=photobiaspars, photopars=photopars,
            IApars=IApars, spectrononlinearpars=spectrononlinearpars, spectrobiaspars=spectrobiaspars,
            IMbiaspars=IMbiaspars, PShotpars=PShotpars, surveyName=surveyName, cosmoModel=cosmoModel,
            latexnames=latexnames)
        self.config = cfg
        self.settings = cfg.settings  # the Config instance is private to this object: no second copy needed
        self.specs = cfg.specs
        self.fiducialcosmopars = dict(cfg.fiducialparams)
        self.fiducialparams = self.fiducialcosmopars
        self.fiducialcosmo = cfg.fiducialcosmo
        self.photopars = dict(cfg.photoparams)
        self.photobiaspars = dict(cfg.Photobiasparams)
        self.IApars = dict(cfg.IAparams)
        self.Spectrobiaspars = dict(cfg.Spectrobiasparams)
        self.Spectrobiasparams = self.Spectrobiaspars
        self.Spectrononlinpars = dict(cfg.Spectrononlinearparams)
        self.Spectrononlinearparams = self.Spectrononlinpars
        self.IMbiaspars = dict(cfg.IMbiasparams)
        self.IMbiasparams = self.IMbiaspars
        self.PShotpars = dict(cfg.PShotparams)
        self.PShotparams = self.PShotpars
        self.observables = list(cfg.obs)
        self.obs = self.observables
        self.input_type = cfg.input_type
        self.external = cfg.external
        self.freeparams = dict(cfg.freeparams)
        self.latex_names = cfg.latex_names
        allpars = {}
        for d in (self.fiducialcosmopars, self.photobiaspars, self.photopars, self.IApars, self.Spectrobiaspars,
                  self.Spectrononlinpars, self.IMbiaspars, self.PShotpars):
            allpars.update(d)
        self.allparams = allpars
        self.allparams_fidus = dict(allpars)
        self.parallel = parallel  # the reference only starts/stops a Ray session here; nothing to do
        self.feed_lvl = self.settings["feedback"]
        self.stencil = self.settings.get("stencil", "3PT")
        self._ctx = None  # created on first use so that configuration / host prep work without a GPU
        self._spectro_job = None
        self._lazy = {}
        self.timings = {}
        self._shard = None  # (rank, world) set by distributed.shard(); default: torch.distributed if initialised

    @property
    def ctx(self) -> _lib.Context:
        if self._ctx is None:
            self._ctx = _lib.default_context(self.settings.get("device"))
        return self._ctx

    # ------------------------------------------------------------------------------------------
    def compute(self, max_z_bins=None):
        """Compute the Fisher matrix of the configured observables; returns a `fisher_matrix` object
        (re-read from the exported text file, like the reference) or None when `outroot` is empty."""
        return self._finish(self._launch(max_z_bins))

    def _launch(self, max_z_bins=None):
        """Host preparation, H2D and the (asynchronous) kernel launches of this forecast; `_finish` collects."""
        t0 = time()
        if "GCph" in self.observables or "WL" in self.observables:
            state = ("photo", self._launch_photo())
        elif "GCsp" in self.observables:
            state = ("spectro", self._launch_spectro(max_z_bins))
        else:
            raise AttributeError("Observables list not defined properly")
        return (t0,) + state

    def _finish(self, launched):
        t0, kind, st = launched
        finalFisher = self._finish_photo(st) if kind == "photo" else self._finish_spectro(st)
        t1 = time()
        self.timings["compute_s"] = t1 - t0
        self.fisher_array = finalFisher
        return self.export_fisher(finalFisher, totaltime=t1 - t0)

    @staticmethod
    def compute_many(forecasts, max_z_bins=None):
        """A batch of forecasts (FisherMatrix objects, any mix of probes) as a two-stage pipeline: the kernels of a
        forecast run asynchronously while the host prepares and uploads the next one; results are collected one
        forecast late.  Returns the list of results in order -- the same objects `compute()` returns.
        (The reference has no batch entry: every forecast is one blocking `compute()`; SURVEY 8e names batches of
        forecasts -- many fiducials or survey specifications -- as the unit that fills a GPU.)"""
        global _deferred
        out, pending = [], None

        def collect(f, launched):
            out.append(f._finish(launched))
            f.release_device()  # a batch keeps at most two forecasts' buffers on the device

        sync_export = os.environ.get("CFP_SYNC_EXPORT", "0") not in ("", "0")  # (A/B switch)
        writer = _DeferredWrites() if _deferred is None and not sync_export else None  # (result files: see _DeferredWrites)
        if writer is not None:
            _deferred = writer
        try:
            for f in forecasts:
                launched = f._launch(max_z_bins)
                if pending is not None:
                    collect(*pending)
                pending = (f, launched)
            if pending is not None:
                collect(*pending)
            if writer is not None:
                writer.flush()  # every file of the batch is on disk when the results are returned
        finally:
            if writer is not None:
                _deferred = None
                writer.close()
        return out

    def release_device(self):
        """Return this forecast's device buffers (plan, table bank) to the context's pool.  Everything that needs them
        later (`derivs_dict`, `veff_arr`, `photo_derivs`, likelihoods built on this object) re-creates them on demand."""
        job = getattr(getattr(self, "photo_LSS", None), "_job", None) or self._spectro_job
        if job is not None:
            job.close()
            cs = getattr(job, "cosmo_set", None)
            if cs is not None:
                cs.release_bank()

    # --- spectro ------------------------------------------------------------------------------
    def set_pk_settings(self):
        """k, mu lattice in 1/Mpc (cosmicfish.py:366-381)."""
        h = self.fiducialcosmopars["h"]
        self.kmin_fish = self.specs["kmin_GCsp"] * h
        self.kmax_fish = self.specs["kmax_GCsp"] * h
        self.Pk_ksamp = self.settings["spectro_Pk_k_samples"]
        self.Pk_musamp = self.settings["spectro_Pk_mu_samples"]
        self.Pk_kgrid = np.linspace(self.kmin_fish, self.kmax_fish, self.Pk_ksamp)
        self.Pk_mugrid = np.linspace(0.0, 1.0, self.Pk_musamp)
        # read-only broadcast views (the fine lattice is 8 MB per mesh): same values and shapes as np.meshgrid's copies
        self.Pk_kmesh, self.Pk_mumesh = np.meshgrid(self.Pk_kgrid, self.Pk_mugrid, copy=False)

    def eliminate_zbinned_freepars(self, nmax):
        """Drop `name_<bin>` parameters of bins beyond nmax (cosmicfish.py:813-839)."""
        for key in list(self.freeparams):
            if "_" in key:
                try:
                    ibin = int(key.split("_")[1])
                except ValueError:
                    continue
                if ibin > nmax:
                    self.freeparams.pop(key)

    def _launch_spectro(self, max_z_bins=None):
        t0 = time()
        self.set_pk_settings()
        self.obs_spectrum = ["g", "g"]
        self.pk_obs_fid = spectro_mod.ComputeGalSpectro(
            cosmopars=self.fiducialcosmopars, fiducial_cosmopars=self.fiducialcosmopars,
            spectrobiaspars=self.Spectrobiaspars, spectrononlinearpars=self.Spectrononlinpars,
            IMbiaspars=self.IMbiaspars, PShotpars=self.PShotpars, configuration=self)
        self.pk_cov = spectro_mod.SpectroCov(self.fiducialcosmopars, fiducial_specobs=self.pk_obs_fid,
                                             bias_samples=self.obs_spectrum, configuration=self)
        self.zmids = self.pk_cov.inter_z_bin_mids
        if max_z_bins is not None and isinstance(max_z_bins, int):
            self.eliminate_zbinned_freepars(max_z_bins)
            self.zmids = self.zmids[0:max_z_bins]
        self.fish_z_arr = np.array(self.zmids)
        self.pk_deriv = spectro_mod.SpectroDerivs(
            self.fish_z_arr, self.Pk_kmesh, self.Pk_mumesh, fiducial_spectro_obj=self.pk_obs_fid,
            bias_samples=self.obs_spectrum, configuration=self, stencil=self.stencil)
        job = self.pk_deriv.build_job(self.pk_cov, freeparams=self.freeparams)
        self._spectro_job = job
        t1 = time()
        rank, world = self._shard if self._shard is not None else cdist.rank_world()
        plan = job.plan()
        plan.set_slice(*cdist.split_range(self.Pk_ksamp * self.Pk_musamp, rank, world))
        plan.run(materialize=0)
        return plan, world, t0, t1

    def _finish_spectro(self, st):
        plan, world, t0, t1 = st
        # partial Fisher matrices of the slices: one all-reduce on the plan's own device buffer (NCCL), then one D2H
        reduced = world > 1 and cdist.allreduce_plan_result(plan, (plan.Z, plan.npar, plan.npar))
        Fz, _, _, _ = plan.fetch()
        if world > 1 and not reduced:
            Fz = cdist.allreduce_numpy(Fz, device=getattr(getattr(plan, "ctx", None), "device", None))  # (gloo group, or no process group at all: no-op)
        plan.set_slice(0, self.Pk_ksamp * self.Pk_musamp)
        t2 = time()
        self.timings.update(host_prep_s=t1 - t0, device_s=t2 - t1,
                            kernel_ms=getattr(getattr(plan, "ctx", None), "last_run_ms", float("nan")))
        self.allparams = self.pk_deriv.fiducial_allpars
        self.parslist = list(self.freeparams.keys())
        self.pk_fisher_per_bin = Fz
        self._lazy.clear()
        return np.sum(Fz, axis=0)

    def _compute_spectro(self, max_z_bins=None):
        return self._finish_spectro(self._launch_spectro(max_z_bins))

    def _materialise_spectro(self):
        if "derivs" not in self._lazy:
            if self._spectro_job is None:
                raise AttributeError("compute() has not been run for a spectroscopic observable")
            derivs = self.pk_deriv.compute_derivs()
            self._lazy["derivs"] = derivs
            self._lazy["veff"] = [self.pk_deriv.veff[i] for i in range(len(self.zmids))]

    @property
    def derivs_dict(self):
        self._materialise_spectro()
        return self._lazy["derivs"]

    @property
    def veff_arr(self):
        self._materialise_spectro()
        return self._lazy["veff"]

    def compute_binned_derivs(self, z_arr):
        """{par: {"z_bins": z_arr, i: d ln P_obs / d par on the (mu, k) lattice}} for the given bin centres
        (cosmicfish.py:383-405): the lattice kernel in materialising mode."""
        from . import spectro as spectro_mod

        self.pk_deriv = spectro_mod.SpectroDerivs(
            z_arr, self.Pk_kmesh, self.Pk_mumesh, fiducial_spectro_obj=self.pk_obs_fid,
            bias_samples=getattr(self, "obs_spectrum", ["g", "g"]),
            configuration=self, stencil=self.stencil)
        return self.pk_deriv.compute_derivs(freeparams=self.freeparams)

    def fish_integrand(self, zi, k, mu, pi, pj):
        """k^2 2 V_survey V_eff dlnP/dpi dlnP/dpj on the lattice (cosmicfish.py:514-542), from the materialised
        derivative and effective-volume lattices of the device run."""
        return k**2 * 2 * self.pk_cov.volume_survey(zi) * self.veff_arr[zi] * self.derivs_dict[pi][zi] * self.derivs_dict[pj][zi]

    def fisher_integral(self, integrand):
        """Simpson over mu, then over k, of a user-supplied lattice (cosmicfish.py:495-512).  A convenience for
        inspecting single elements; compute() never calls it -- the fused kernel does the same sum with the same
        weights on the device."""
        from .quadrature import simpson_weights

        integrand = np.asarray(integrand, dtype=float)
        return float(simpson_weights(self.Pk_kgrid) @ (simpson_weights(self.Pk_mugrid) @ integrand))

    def fisher_calculation(self, zi, p_i, p_j):
        """One element of the Fisher matrix of bin zi (cosmicfish.py:472-493)."""
        return self.fisher_integral(self.fish_integrand(zi, self.Pk_kmesh, self.Pk_mumesh, p_i, p_j))

    def recap_options(self):
        """Print the selected options (cosmicfish.py:1004-1050)."""
        if self.settings["feedback"] > 1:
            print("\n----------RECAP OF SELECTED OPTIONS--------\n\nSettings:")
            for key in self.settings:
                print("   " + key + ": {}".format(self.settings[key]))
            print("\nSpecifications:")
            for key in self.specs:
                print("   " + key + ": {}".format(self.specs[key]))
            for title, d in (("Cosmological parameters:", self.fiducialcosmopars), ("Photometric parameters:", self.photopars),
                             ("IA parameters:", self.IApars), ("Bias parameters:", self.photobiaspars),
                             ("SpectroBias parameters:", self.Spectrobiaspars), ("Shot noise parameters:", self.PShotpars)):
                print(title)
                for key in d:
                    print("   " + key + ": {}".format(d[key]))

    def pk_LSS_Fisher(self, nbins):
        """Per-bin Fisher matrices [nbins][npar][npar] (cosmicfish.py:407-443), from the fused kernel."""
        return np.array(self.pk_fisher_per_bin[:nbins])

    def fisher_per_bin(self, ibin):
        return np.array(self.pk_fisher_per_bin[ibin])

    # --- photo --------------------------------------------------------------------------------
    def _launch_photo(self):
        from . import photo as photo_mod

        t0 = time()
        self.photo_obs_fid = photo_mod.ComputeCls(self.fiducialcosmopars, self.photopars, self.IApars,
                                                  self.photobiaspars, print_info_specs=True,
                                                  fiducial_cosmo=self.fiducialcosmo, configuration=self)
        self.photo_LSS = photo_mod.PhotoCov(self.fiducialcosmopars, self.photopars, self.IApars, self.photobiaspars,
                                            fiducial_Cls=self.photo_obs_fid, configuration=self,
                                            stencil=self.stencil)
        t1 = time()
        shard = self._shard if self._shard is not None else cdist.rank_world()
        return self.photo_LSS.launch_fisher(self.freeparams, shard=shard), t0, t1

    def _finish_photo(self, st):
        launched, t0, t1 = st
        F = self.photo_LSS.finish_fisher(launched)
        t2 = time()
        plan = self.photo_LSS._job.plan()
        self.timings.update(host_prep_s=self.photo_LSS.timings.get("host_prep_s", t1 - t0),
                            device_s=self.photo_LSS.timings.get("device_s", t2 - t1),
                            kernel_ms=getattr(getattr(plan, "ctx", None), "last_run_ms", float("nan")))
        return F

    def _compute_photo(self):
        return self._finish_photo(self._launch_photo())

    @property
    def photo_derivs(self):
        return self.photo_LSS.compute_derivs()

    def photo_LSS_fishermatrix_einsum(self, noisy_cls=None, covmat=None, derivs=None, lss_obj=None):
        """The Fisher matrix of a photometric probe (cosmicfish.py:544-744).

        Called without arrays (or with the very objects `lss_obj.compute_covmat()` / `compute_derivs()` returned) the
        configured job is assembled in one device pass from its own stencil.  Called with the caller's own arrays --
        `noisy_cls` (for its "ells"), `covmat` (one matrix or DataFrame per multipole) and `derivs`
        ({parameter: {"X ixY j": dC/dtheta (Nl,)}}, from any derivative scheme) -- the reference's assembly
        sum_l (l_c + 1/2) dl Tr(dC_p S^-1 dC_q S^-1) with left-edge arrays runs on the device on exactly those arrays
        (cfp_photo_fisher)."""
        obj = lss_obj if lss_obj is not None else getattr(self, "photo_LSS", None)
        own = getattr(obj, "__dict__", {})
        given = {"noisy_cls": noisy_cls, "covmat": covmat, "derivs": derivs}
        if all(v is None or v is own.get(k) for k, v in given.items()):
            return obj.compute_fisher(self.freeparams)
        if covmat is None and obj is not None:
            noisy_cls, covmat = obj.compute_covmat()
        if derivs is None and obj is not None:
            derivs = obj.compute_derivs()
        if noisy_cls is None or covmat is None or derivs is None:
            raise ValueError("photo_LSS_fishermatrix: pass noisy_cls, covmat and derivs (or an lss_obj to take them from)")
        cols = []
        for o in self.observables:  # cosmicfish.py:686-691
            if o in ("GCph", "WL"):
                cols.extend(f"{o} {i + 1}" for i in range(len(self.specs["z_bins_" + o]) - 1))
        cov = np.array([np.asarray(c, dtype=float) for c in covmat])
        dC = np.array([[[np.asarray(derivs[par][f"{a}x{b}"], dtype=float) for b in cols] for a in cols]
                       for par in self.freeparams])                      # [npar][Nb][Nb][Nl]
        dC = np.ascontiguousarray(np.transpose(dC, (0, 3, 1, 2)))        # [npar][Nl][Nb][Nb]
        return self.ctx.photo_fisher(cov, dC, np.asarray(noisy_cls["ells"], dtype=float))

    photo_LSS_fishermatrix = photo_LSS_fishermatrix_einsum

    # --- export -------------------------------------------------------------------------------
    _git_version_cache = {}

    @classmethod
    def _git_version(cls):
        cwd = os.getcwd()  # one `git rev-parse` per working directory, not one subprocess per Fisher matrix
        if cwd not in cls._git_version_cache:
            try:
                v = subprocess.check_output(["git", "rev-parse", "HEAD"], stderr=subprocess.DEVNULL).decode().strip()
            except Exception:
                v = "unknown"
            cls._git_version_cache[cwd] = v
        return cls._git_version_cache[cwd]

    _spec_text_cache = {}

    def _spec_texts(self):
        """(sections of `_specs.dat`, the `_specs.json` object after its metadata entry) of this forecast's settings,
        fiducial, free parameters and specifications.  Remembered by CONTENT (a pickle of the dictionaries is the key:
        40 us against 0.3 ms of formatting), so a series of forecasts that share them formats them once."""
        try:
            key = pickle.dumps((self.settings, self.specs, self.fiducialcosmopars, self.freeparams, self.allparams), 5)
        except Exception:
            key = None
        hit = self._spec_text_cache.get(key) if key is not None else None
        if hit is not None:
            return hit
        lines = []
        for title, d in (("CosmicFish settings", self.settings), ("Fiducial parameters", self.fiducialcosmopars),
                         ("Free parameters", self.freeparams), ("Survey specifications", self.specs)):
            lines.append("### %s ###\n" % title)
            for k, value in sorted(d.items()):
                lines.append("%s:%s\n" % (k, _array_str(value) if isinstance(value, np.ndarray) else value))
        try:
            body = json.dumps({"options": _jsonable(self.settings), "specifications": _jsonable(self.specs),
                               "fiducialpars": _jsonable(self.fiducialcosmopars),
                               "freepars": _jsonable(self.freeparams), "allparams": _jsonable(self.allparams)},
                              default=_json_default)[1:]
            # (no indent: the C-accelerated encoder only handles the compact form; same content as the reference's file)
        except Exception:
            body = None
        hit = ("".join(lines), body)
        if key is not None:
            if len(self._spec_text_cache) >= 64:
                self._spec_text_cache.clear()
            self._spec_text_cache[key] = hit
        return hit

    def export_fisher(self, fishmat, totaltime=None):
        """Write <results_dir>/<cf_version>_<outroot>_<obs>_FM{.txt,.paramnames,_specs.dat,_specs.json} and
        return the matrix re-read from the text file (cosmicfish.py:841-1002); None if outroot == ''."""
        if self.settings["outroot"] == "":
            return None
        obstring = "".join(self.observables)
        cols = list(self.freeparams)
        os.makedirs(self.settings["results_dir"], exist_ok=True)
        root = self.settings["results_dir"] + "/" + self.cf_version + "_" + self.settings["outroot"] + "_" + obstring + "_FM"
        # Under torchrun every rank holds the same reduced matrix and runs this method.  Files are written to a
        # rank-private temporary name and moved into place atomically, so a rank can never read another rank's
        # half-written file; the result object of every rank is built from the text it wrote itself (repr round-trip,
        # exactly what re-reading the file gives).
        rank = (self._shard if self._shard is not None else cdist.rank_world())[0]
        if cdist.is_distributed():
            tag = ".tmp%d_%d" % (rank, os.getpid())

            def write_atomic(path, text):
                _write_now(path + tag, text)
                os.replace(path + tag, path)
        else:  # a single process owns its files: plain writes
            write_atomic = _write_text

        arr = np.atleast_2d(np.asarray(fishmat, dtype=float))
        fm_text = "#" + " ".join(["", *cols]) + "\n" + "".join("\t".join(map(repr, row)) + "\n" for row in arr.tolist())
        fid6 = ["{:.6f}".format(self.allparams[par]) for par in cols]
        latex = [str(self.latex_names.get(par, str(par))) for par in cols]
        pn_text = "#\n#\n# This file contains the parameter names for a derived Fisher matrix.\n#\n" + "".join(
            str(par) + "    " + lt + "    " + f6 + "\n" for par, lt, f6 in zip(cols, latex, fid6))
        write_atomic(root + self.settings["fishermatrix_file_extension"], fm_text)
        write_atomic(root + ".paramnames", pn_text)
        if arr.any(axis=1).all() and arr.any(axis=0).all():
            # what re-reading the two files gives (repr round-trips exactly, fiducials carry six decimals), without
            # parsing them back
            out = fm.fisher_matrix(fisher_matrix=arr, param_names=cols, param_names_latex=latex,
                                   fiducial=[float(f6) for f6 in fid6])
            out.file_name = root + ".txt"
            out.path = os.path.abspath(out.file_name)
            out.name = os.path.splitext(os.path.basename(out.file_name))[0]
            out.indir = os.path.dirname(out.path)
        else:  # all-zero rows / columns are deleted on reading (analysis/fisher_matrix.py:261-268): take that route
            if _deferred is not None:
                _deferred.flush()
            out = fm.fisher_matrix(file_name=root + ".txt")
        if rank != 0 and cdist.is_distributed():
            return out  # the specification files are rank 0's business
        git = self._git_version()
        head = "**Git version commit: %s \n" % git
        if totaltime is not None:
            head += "**Time needed for computation of Fisher matrix: {:.3f} seconds \n".format(totaltime)
        head += "**Observables: %s \n" % obstring
        dat_body, json_body = self._spec_texts()
        _write_text(root + "_specs.dat", head + dat_body)
        if json_body is not None:  # best effort, like the reference
            meta = json.dumps({"cf_version": self.cf_version, "git_commit": git, "observables": list(self.observables),
                               "totaltime_sec": None if totaltime is None else float(totaltime),
                               "backend": "cosmicfishpie_b200"})
            _write_text(root + "_specs.json", '{"metadata": ' + meta + ", " + json_body)
        return out

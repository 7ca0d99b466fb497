"""Nautilus driver for the batched GPU likelihoods (reference: likelihood/sampler.py:14-261, likelihood/base.py:176-220).

Same configuration dictionary as the reference's `NautilusSampler` -- `name`, `fiducial`, `observables`, `options`,
`priors` (ranges or {"type": "gaussian", "loc", "scale"}), `sampler_settings` (`n_live`, `n_networks`, `n_batch`,
`pool`), `use_nuisance` -- and the same outputs: `chains/chains_<name>/cosmicjellyfish_<code>_<survey>_<name>.txt`
(columns: parameters, weights, log-likelihood; header as the reference writes it), `<outroot>_metadata.json` with the
reference's keys, and the sampler's own checkpoint `<outroot>.hdf5`.  What differs is how the likelihood is called:

* the reference hands Nautilus the scalar `loglike` and a process pool (`pool=N`: N forked CPU workers, one point
  each).  Forking after CUDA initialisation is not an option and one point per call wastes the GPU, so this driver
  passes the VECTORISED entry (`vectorized=True, pool=None`): every batch Nautilus proposes (n_batch points) goes to
  `loglike_batch` as one matrix and, under torch.distributed, is sharded over the GPUs with one all-gather;
* the reference wires only the photometric likelihood (sampler.py:139-142).  Here `observables` may name photometric
  probes, "GCsp", or both: both means the JOINT likelihood, the sum of the two log-likelihoods on one parameter vector
  (each probe ignores the parameters it does not know, like the reference's compute_theory);
* this package consumes cached Boltzmann tables, so the configuration also carries `extfiles` (the table bank).
`nautilus` is imported lazily: it is an optional dependency here as in the reference."""
from __future__ import annotations

import json
import os
import time
from datetime import datetime

import numpy as np

_LABELS = {"Omegam": r"$\Omega_m$", "Omegab": r"$\Omega_b$", "Omegac": r"$\Omega_c$", "Omegak": r"$\Omega_k$",
           "sigma8": r"$\sigma_8$", "ns": r"$n_{\rm s}$", "w0": r"$w_0$", "wa": r"$w_a$", "h": r"$h$",
           "A_s": r"$10^9 A_s$", "As": r"$10^9 A_s$", "H0": r"$H_0$", "mnu": r"$m_\nu$", "Neff": r"$N_{\rm eff}$",
           "AIA": r"$A_{\rm IA}$", "etaIA": r"$\eta_{\rm IA}$"}


def _format_param_label(name: str) -> str:
    """LaTeX label of a parameter (sampler.py:14-43: Omega*/omega* prefixes and b<i> first, then the table)."""
    for prefix, sym in (("Omega", r"\Omega"), ("omega", r"\omega")):
        if name.startswith(prefix):
            rest = name[len(prefix):]
            return rf"${sym}_{{{rest}}}$" if rest else rf"${sym}$"
    if name[:1] == "b" and name[1:].isdigit():
        return rf"$b_{{{name[1:]}}}$"
    return _LABELS.get(name, name)


def load_chain_metadata(chain_folder, metadata_filename=None, label_overrides=None):
    """(chain file, sampled fiducial values, metadata, labels) of a finished run (sampler.py:46-83)."""
    if metadata_filename is None:
        found = [f for f in os.listdir(chain_folder) if f.endswith("_metadata.json")]
        if len(found) != 1:
            raise ValueError(f"Expected exactly one metadata file in {chain_folder}, found: {found}")
        metadata_filename = found[0]
    path = os.path.join(chain_folder, metadata_filename)
    with open(path) as fh:
        meta = json.load(fh)
    chain = meta.get("chain_file")
    if chain is None and meta.get("outroot path"):
        chain = meta["outroot path"] + ".txt"
    if chain is None:
        raise ValueError(f"Chain file not found in metadata: {path}")
    if not os.path.isabs(chain):
        chain = os.path.join(chain_folder, os.path.basename(chain))
    fid = meta.get("sampled_fiducial_params", {})
    over = label_overrides or {}
    return chain, fid, meta, {k: over.get(k, _format_param_label(k)) for k in fid}


class JointLikelihood:
    """Sum of independent log-likelihoods evaluated on the same parameter vector (photometric + spectroscopic)."""

    def __init__(self, parts):
        self.parts = list(parts)

    def loglike(self, param_vec=None, *, param_dict=None, prior=None) -> float:
        return float(sum(p.loglike(param_vec, param_dict=param_dict, prior=prior) for p in self.parts))

    def loglike_batch(self, param_matrix, prior=None, keys=None, shard=False) -> np.ndarray:
        return sum(p.loglike_batch(param_matrix, prior=prior, keys=keys, shard=shard) for p in self.parts)


class NautilusSampler:
    def __init__(self, config, base_dir="."):
        from .fishermatrix import FisherMatrix
        from .likelihood import PhotometricLikelihood, SpectroLikelihood

        self.config = config
        self.fiducial = config["fiducial"]
        self.observables = list(config["observables"])
        self.options = config["options"]
        self.prior_dict = config["priors"]
        self.sampler_settings = config["sampler_settings"]
        self.use_nuisance = config.get("use_nuisance", True)
        self.tini = self.tfin = None
        self.timestamp = datetime.now().strftime("%Y_%m_%d_%H_%M")
        self.folder_name = os.path.join(base_dir, "chains", f"chains_{config['name']}")
        os.makedirs(self.folder_name, exist_ok=True)
        self.outroot = os.path.join(self.folder_name,
                                    f"cosmicjellyfish_{self.options['code']}_{self._get_survey_name()}_{config['name']}")
        if self.options.get("feedback", 0):
            print(f"NautilusSampler '{config['name']}': output {self.outroot}, pool setting of the configuration "
                  f"{self.sampler_settings.get('pool')} (ignored: batches go to the GPU)")
        # one FisherMatrix (fiducial cosmology, data vector) per probe family
        photo = [o for o in self.observables if o in ("WL", "GCph")]
        spectro = [o for o in self.observables if o == "GCsp"]
        if not photo and not spectro:
            raise ValueError("observables must name WL / GCph and / or GCsp")
        self.fishers, parts = {}, []
        for obs, cls in ((photo, PhotometricLikelihood), (spectro, SpectroLikelihood)):
            if not obs:
                continue
            opts = dict(self.options)
            opts["outroot"] = os.path.basename(self.outroot) + "_" + "".join(obs)
            opts.setdefault("results_dir", self.folder_name)
            if obs == spectro:
                opts["survey_name_photo"] = False
            else:
                opts["survey_name_spectro"] = False
            fm = FisherMatrix(fiducialpars=dict(self.fiducial), options=opts, observables=list(obs),
                              extfiles=config["extfiles"], cosmoModel=self.options["cosmo_model"],
                              surveyName=self.options["survey_name"], freepars=config.get("freepars"))
            if obs == spectro:
                fm.compute()  # the spectroscopic likelihood reads the covariance objects compute() builds
                parts.append(cls(cosmoFM_data=fm, cosmoFM_theory=fm, leg_flag=config.get("leg_flag", "wedges")))
            else:
                parts.append(cls(cosmo_data=fm, cosmo_theory=fm))
            self.fishers["".join(obs)] = fm
        self.cosmoFM_fid = next(iter(self.fishers.values()))
        self.likelihood = parts[0] if len(parts) == 1 else JointLikelihood(parts)
        self.photo_like = self.likelihood  # the reference's attribute name
        self._setup_priors()

    def _get_survey_name(self):
        return self.options.get("survey_name_photo") or self.options.get("survey_name_spectro")

    def _all_params(self):
        out = {}
        for fm in self.fishers.values():
            out.update(fm.allparams)
        return out

    def _free_params(self):
        out = {}
        for fm in self.fishers.values():
            out.update(fm.freeparams)
        return out

    def _setup_priors(self):
        """Parameters of `priors` that are free parameters of a configured probe, ranges as tuples, gaussians as
        frozen scipy distributions (sampler.py:129-137); nuisance parameters are dropped when use_nuisance is false."""
        from nautilus import Prior
        from scipy.stats import norm

        free = self._free_params()
        cosmo = set(self.cosmoFM_fid.fiducialcosmopars)
        self.prior_chosen = Prior()
        for par, rng in self.prior_dict.items():
            if par not in free or (not self.use_nuisance and par not in cosmo):
                continue
            if isinstance(rng, dict) and rng.get("type") == "gaussian":
                dist = norm(loc=rng["loc"], scale=rng["scale"])
            else:
                dist = tuple(rng)
            self.prior_chosen.add_parameter(par, dist)

    def _save_metadata(self, evidence=None, finish_time=None):
        allp = self._all_params()
        meta = {"cosmicfishpie_version": "1.0.0", "backend": "cosmicfishpie_b200", "name": self.config["name"],
                "start_time": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(self.tini)),
                "outroot path": self.outroot, "cosmo_fiducial_params": self.fiducial,
                "sampled_fiducial_params": {p: allp[p] for p in self.prior_dict if p in allp},
                "priors": {k: str(v) for k, v in self.prior_dict.items()},
                "sampler_settings": self.sampler_settings, "observables": self.observables,
                "cosmo_model": self.options["cosmo_model"], "survey_name": self._get_survey_name(),
                "code": self.options["code"]}
        if finish_time:
            meta.update({"finish_time": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(finish_time)),
                         "elapsed_time": self._format_time(finish_time - self.tini),
                         "evidence_log_z": float(evidence) if evidence else None, "chain_file": self.chain_file})
        with open(self.outroot + "_metadata.json", "w") as fh:
            json.dump(meta, fh, indent=2)

    @staticmethod
    def _format_time(seconds):
        s = int(seconds)
        return f"{s // 3600:02d}:{(s % 3600) // 60:02d}:{s % 60:02d}"

    def _make_sampler(self):
        from nautilus import Sampler

        st = self.sampler_settings
        prior = self.prior_chosen
        return Sampler(prior=prior, likelihood=lambda x: self.likelihood.loglike_batch(x, prior=prior, shard=True),
                       n_live=st["n_live"], n_networks=st["n_networks"], n_batch=st["n_batch"], pool=None,
                       pass_dict=False, vectorized=True, filepath=self.outroot + ".hdf5", resume=True)

    def run(self):
        self.chain_file = self.outroot + ".txt"
        self.chain_hdf5 = self.outroot + ".hdf5"
        self.tini = time.time()
        self._save_metadata()
        if os.path.exists(self.chain_file):  # finished run: refresh the metadata only (sampler.py:208-226)
            evidence = None
            if os.path.exists(self.chain_hdf5):
                try:
                    evidence = self._make_sampler().evidence()
                except Exception as exc:  # an unreadable checkpoint does not invalidate the chain
                    print(f"WARNING: Could not load existing run: {exc}")
            self._save_metadata(evidence=evidence, finish_time=time.time())
            return None
        sampler = self._make_sampler()
        sampler.run(verbose=bool(self.options.get("feedback", 0)), discard_exploration=True)
        evidence = sampler.log_z
        points, log_w, log_l = sampler.posterior()
        self.tfin = time.time()
        chain = np.column_stack([points, np.exp(log_w), log_l])
        np.savetxt(self.chain_file, chain, header="loglike weights " + " ".join(self.prior_chosen.keys))
        self._save_metadata(evidence=evidence, finish_time=self.tfin)
        return sampler

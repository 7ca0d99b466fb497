"""Run configuration: options, survey specifications, parameter dictionaries.

Honours the same keys and precedence as the reference's `cfg.init`
(cosmicfishpie/configs/config.py:18-870) for the in-scope probes (GCsp, WL, GCph), so the same
`options / specifications / freepars / extfiles / fiducialpars` dictionaries drive both codes:

  settings defaults ............ config.py:204-280
  external-file layout ......... config.py:286-340
  survey YAML loading .......... config.py:376-590  (defaults first, then the named survey, then
                                                     user `specifications`)
  default free/fiducial pars ... config.py:597-646
  nuisance dictionaries ........ config.py:665-796  (free parameters are appended in this order:
                                                     cosmology, photo bias, IA, spectro bias, P_shot,
                                                     non-linear sigma_p/sigma_v)

Unlike the reference this is an object (`Config`), not mutable module state, so several
FisherMatrix instances can coexist.  `current()` returns the most recently initialised one for
code written against the module-global style.
"""
from __future__ import annotations

import copy
import os
import pickle
from typing import Optional

import numpy as np
import yaml

SR = np.power(180 / np.pi, 2)
AREASKY = 4 * np.pi * SR

_DATA = os.path.join(os.path.dirname(os.path.realpath(__file__)), "data")
SPECS_DIR_DEFAULT = os.path.join(_DATA, "specs")
DEFAULT_SPECTRO = "Euclid-Spectroscopic-ISTF-Pessimistic"
DEFAULT_PHOTO = "Euclid-Photometric-ISTF-Pessimistic"

_SETTING_DEFAULTS = {
    "survey_specs": "ISTF-Optimistic",
    "survey_name_photo": DEFAULT_PHOTO,
    "survey_name_spectro": DEFAULT_SPECTRO,
    "fail_on_specs_not_found": False,
    "derivatives": "3PT",
    "nonlinear": True,
    "nonlinear_photo": True,
    "bfs8terms": False,
    "vary_bias_str": "lnb",
    "AP_effect": True,
    "FoG_switch": True,
    "GCsp_linear": False,
    "fix_cosmo_nl_terms": True,
    "Pshot_nuisance_fiducial": 0.0,
    "pivot_z_IA": 0.0,
    "accuracy": 1,
    "spectro_Pk_k_samples": 1025,
    "spectro_Pk_mu_samples": 17,
    "feedback": 2,
    "activateMG": False,
    "external_activateMG": False,
    "outroot": "default_run",
    "code": "external",
    "memorize_cosmo": False,
    "results_dir": "./results",
    "SUPPRESS_WARNINGS": True,
    "fishermatrix_file_extension": ".txt",
    "savgol_window": 101,
    "savgol_polyorder": 3,
    "savgol_width": 1.358528901113328,
    "savgol_internalsamples": 800,
    "savgol_internalkmin": 0.001,
    "eps_cosmopars": 0.01,
    "eps_gal_nuispars": 0.0001,
    "eps_gal_nonlinpars": 0.01,
    "GCsp_Tracer": "matter",
    "GCph_Tracer": "matter",
    "ell_sampling": "accuracy",
    "ShareDeltaNeff": False,
    "kh_rescaling_bug": False,
    "kh_rescaling_beforespecerr_bug": False,
}

_EXT_DEFAULTS = {
    "file_names": {
        "H_z": "background_Hz", "D_zk": "D_Growth-zk", "f_zk": "f_GrowthRate-zk", "fcb_zk": "fcb_GrowthRate-zk",
        "Pl_zk": "Plin-zk", "Plcb_zk": "Plincb-zk", "Pnl_zk": "Pnonlin-zk", "Pnlcb_zk": "Pnonlincb-zk",
        "s8_z": "sigma8-z", "s8cb_z": "sigma8cb-z", "SigmaWL": None, "z_arr": "z_values_list",
        "k_arr": "k_values_list", "k_arr_special": None,
    },
    "E-00": True,
    "fiducial_folder": "fiducial_eps_0",
    "folder_paramnames": [],
}

_FIDUCIAL_DEFAULT = {
    "Omegam": 0.32, "Omegab": 0.05, "w0": -1.0, "wa": 0.0, "h": 0.67, "ns": 0.96, "sigma8": 0.815583,
    "Omegak": 0.0, "mnu": 0.06, "tau": 0.058, "num_nu_massive": 1, "num_nu_massless": 2.046,
    "dark_energy_model": "ppf",
}

_LATEX = {
    "Omegam": r"\Omega_{{\rm m}, 0}", "Omegab": r"\Omega_{{\rm b}, 0}", "h": r"h", "ns": r"n_{\rm s}",
    "sigma8": r"\sigma_8", "w0": r"w_0", "wa": r"w_a", "AIA": r"A_{IA}", "betaIA": r"\beta_{IA}",
    "etaIA": r"\eta_{IA}", "mnu": r"\Sigma m_\nu", "tau": r"\tau",
}
for _i in range(1, 14):
    _LATEX[f"lnbg_{_i}"] = rf"\ln(b_g)_{{{_i}}}"
    _LATEX[f"bg_{_i}"] = rf"b_{{g{_i}}}"
    _LATEX[f"Ps_{_i}"] = rf"P_{{S{_i}}}"
    _LATEX[f"b{_i}"] = rf"b_{{{_i}}}"
    _LATEX[f"sigmap_{_i}"] = rf"\sigma_{{p\,{_i}}}"
    _LATEX[f"sigmav_{_i}"] = rf"\sigma_{{v\,{_i}}}"


def _deepupdate(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deepupdate(dst[k], v)
        else:
            dst[k] = v


def _ngal_per_bin(ngal_sqarmin, zbins):
    nbins = len(zbins[:-1])
    return ((ngal_sqarmin * 3600) / nbins) * SR * np.ones_like(np.asarray(zbins[:-1], dtype=float))


_YAML_CACHE = {}


def _clone(o):
    """Deep copy of plain YAML data (dict / list / scalars / arrays): what copy.deepcopy does for these types without
    its memo bookkeeping (several times faster; a Config clones four specification files)."""
    if isinstance(o, dict):
        return {k: _clone(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_clone(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.copy()
    if isinstance(o, (str, int, float, bool, type(None), range, tuple)):
        return o
    return copy.deepcopy(o)


def _load_yaml_specs(folder, name, default_name, fail):
    """`specifications:` mapping of a survey YAML; parsed once per (path, mtime), a private deep copy per use (kept
    pickled: unpickling plain YAML data is five times faster than walking it in Python, and a Config takes four)."""
    return pickle.loads(_load_yaml_specs_cached(folder, name, default_name, fail))


def _load_yaml_specs_cached(folder, name, default_name, fail):
    path = os.path.join(folder, name + ".yaml")
    if not os.path.isfile(path):
        if fail:
            raise FileNotFoundError(f"specifications file : {path} not found! Exiting...")
        print(f"WARNING: specifications file : {path} not found!")
        path = os.path.join(SPECS_DIR_DEFAULT, default_name + ".yaml")
        print(f"Using default specifications: {path}")
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _YAML_CACHE:
        with open(path, "r") as fh:
            _YAML_CACHE[key] = pickle.dumps(yaml.safe_load(fh)["specifications"], protocol=pickle.HIGHEST_PROTOCOL)
    return _YAML_CACHE[key]


def _photo_specs(folder, name, fail):
    if not name:
        return {}
    d = _load_yaml_specs(folder, name, DEFAULT_PHOTO, fail)
    d["z_bins_WL"] = np.array(d.get("z_bins_WL", d.get("z_bins_ph")))
    d["ngal_sqarmin_WL"] = d.get("ngal_sqarmin_WL", d.get("ngal_sqarmin"))
    d["ngalbin_WL"] = _ngal_per_bin(d["ngal_sqarmin_WL"], d["z_bins_WL"])
    d["binrange_WL"] = range(1, len(d["z_bins_WL"]))
    d["z_bins_GCph"] = d.get("z_bins_GCph", d.get("z_bins_ph"))
    d["ngal_sqarmin_GCph"] = d.get("ngal_sqarmin_GCph", d.get("ngal_sqarmin"))
    d["ngalbin_GCph"] = _ngal_per_bin(d["ngal_sqarmin_GCph"], d["z_bins_GCph"])
    d["binrange_GCph"] = range(1, len(d["z_bins_GCph"]))
    d["z0"] = d["zm"] / np.sqrt(2)
    d["z0_p"] = d["z0"]
    return d


def _spectro_specs(folder, name, fail):
    if not name:
        return {}
    return _load_yaml_specs(folder, name, DEFAULT_SPECTRO, fail)


class Config:
    """Holds what the reference keeps in `cosmicfishpie.configs.config` module globals."""

    def __init__(self, options=None, specifications=None, observables=None, freepars=None, extfiles=None,
                 fiducialpars=None, photobiaspars=None, photopars=None, IApars=None, PShotpars=None,
                 spectrobiaspars=None, spectrononlinearpars=None, IMbiaspars=None, surveyName="Euclid",
                 cosmoModel="w0waCDM", latexnames=None):
        options = {} if options is None else options
        specifications = {} if specifications is None else specifications
        s = options  # like the reference, defaults are written back into the caller's dict
        s.setdefault("specs_dir_default", SPECS_DIR_DEFAULT)
        s.setdefault("specs_dir", s["specs_dir_default"])
        s.setdefault("external_data_dir", _DATA)
        s.setdefault("survey_name", surveyName)
        s.setdefault("cosmo_model", cosmoModel)
        for k, v in _SETTING_DEFAULTS.items():
            s.setdefault(k, v)
        self.settings = s

        if s["code"] != "external" or extfiles is None:
            raise ValueError(
                "cosmicfishpie_b200 consumes cached Boltzmann tables: pass options['code']='external' and `extfiles` "
                "(the Boltzmann solvers are outside the scope of this package)."
            )
        self.input_type = "external"
        ext = _clone(dict(_EXT_DEFAULTS))
        tables = extfiles.get("tables") if isinstance(extfiles, dict) else None
        _deepupdate(ext, {k: v for k, v in extfiles.items() if k != "tables"})
        ext["tables"] = tables
        if tables is None:
            if not os.path.isdir(ext["directory"]):
                raise ValueError("External directory does not exist")
            if not os.path.isdir(os.path.join(ext["directory"], ext["fiducial_folder"])):
                raise ValueError("External directory does not contain fiducial folder ")
        elif ext["fiducial_folder"] not in tables:
            raise ValueError("In-memory table bank does not contain the fiducial folder")
        self.external = ext

        # --- survey specifications -------------------------------------------------------------
        fail = s["fail_on_specs_not_found"]
        specs = {}
        specs.update(_photo_specs(s["specs_dir_default"], DEFAULT_PHOTO, fail))
        specs.update(_spectro_specs(s["specs_dir_default"], DEFAULT_SPECTRO, fail))
        from_files = {}
        if "Euclid" in surveyName or "DESI" in surveyName or "SKA" in surveyName:
            if s.get("survey_name_spectro"):
                from_files.update(_spectro_specs(s["specs_dir"], s["survey_name_spectro"], fail))
        if "Euclid" in surveyName or "Rubin" in surveyName or "SKA" in surveyName:
            if s.get("survey_name_photo"):
                from_files.update(_photo_specs(s["specs_dir"], s["survey_name_photo"], fail))
        specs.update(from_files)
        for probe in ("GCph", "WL", "spectro", "IM"):
            specs["fsky_" + probe] = from_files.get(
                "fsky_" + probe, from_files.get("area_survey_" + probe, 0.0) / AREASKY)
        specs.update(specifications)
        specs["survey_name"] = surveyName
        specs["specs_dir"] = s["specs_dir"]
        self.specs = specs

        self.obs = ["GCph", "WL"] if observables is None else observables
        obs = self.obs

        if freepars is None:
            e = s["eps_cosmopars"]
            if s["cosmo_model"] == "w0waCDM":
                freepars = {k: e for k in ("Omegam", "Omegab", "w0", "wa", "h", "ns", "sigma8")}
            elif s["cosmo_model"] == "LCDM":
                freepars = {k: e for k in ("Omegam", "Omegab", "h", "ns", "sigma8")}
            else:
                raise ValueError("pass `freepars` explicitly for this cosmological model")
        self.freeparams = _clone(freepars)
        self.fiducialparams = _clone(_FIDUCIAL_DEFAULT if fiducialpars is None else fiducialpars)

        # --- photometric nuisance -----------------------------------------------------------------
        if "GCph" in obs:
            if photobiaspars is None:
                model = specs["ph_bias_model"]
                prtz = specs["ph_bias_parametrization"]
                photobiaspars = {"bias_model": model}
                if model in ("binned", "binned_constant") and prtz[model]["generate_bias_keys"]:
                    zb = specs.get("z_bins_GCph", specs.get("z_bins_ph"))
                    for ind in range(1, len(zb)):
                        photobiaspars[prtz[model]["keystr"] + str(ind)] = np.sqrt(1 + 0.5 * (zb[ind] + zb[ind - 1]))
                else:
                    for k, v in prtz[model].items():
                        if k not in ("generate_bias_keys", "keystr"):
                            photobiaspars[k] = v
        else:
            photobiaspars = {}
        self.Photobiasparams = _clone(photobiaspars)
        if "GCph" in obs and specs.get("vary_ph_bias") is not None:
            for k in photobiaspars:
                if k != "bias_model":
                    self.freeparams.setdefault(k, specs["vary_ph_bias"])
        self.photoparams = _clone(specs.get("photo_z_params") if photopars is None else photopars)
        self.IAparams = _clone(specs.get("IA_params") if IApars is None else IApars)
        if "WL" in obs and specs.get("vary_IA_pars") is not None:
            for k in self.IAparams:
                if k != "IA_model":
                    self.freeparams.setdefault(k, specs["vary_IA_pars"])

        # --- spectroscopic nuisance ---------------------------------------------------------------
        spectro = "GCsp" in obs or "IM" in obs
        self.Spectrononlinearparams = {}
        if spectro and specs.get("nonlinear_model", "default") == "rescale_sigma_pv":
            self.Spectrononlinearparams.update(specs["nonlinear_parametrization"]["rescale_sigma_pv"])
        if spectrononlinearpars is not None:
            self.Spectrononlinearparams.update(spectrononlinearpars)
        self.Spectrobiasparams = copy.deepcopy(spectrobiaspars) if spectrobiaspars is not None else {}
        self.PShotparams = {}
        if PShotpars is not None:
            self.PShotparams = copy.deepcopy(PShotpars)
        elif spectro:  # (sic) the reference fills the bias dictionary only when PShotpars is None
            for k, v in _clone(specs["sp_bias_parametrization"][specs["sp_bias_model"]]).items():
                self.Spectrobiasparams[k] = v
            shot = specs["shot_noise_parametrization"][specs["shot_noise_model"]]
            if isinstance(shot, dict):
                self.PShotparams.update(shot)
        self.IMbiasparams = copy.deepcopy(IMbiaspars) if IMbiaspars is not None else {}
        if "IM" in obs:
            raise NotImplementedError("intensity mapping (IM) is outside the scope of cosmicfishpie_b200")
        if "GCsp" in obs:
            for k in self.Spectrobiasparams:
                self.freeparams.setdefault(k, s["eps_gal_nuispars"])
            for k in self.PShotparams:
                self.freeparams.setdefault(k, s["eps_gal_nuispars"])
            for k in self.Spectrononlinearparams:
                self.freeparams.setdefault(k, s["eps_gal_nonlinpars"])

        self.latex_names = dict(_LATEX)
        if isinstance(latexnames, dict):
            self.latex_names.update(latexnames)

        global _CURRENT
        _CURRENT = self
        from .cosmology import cached_cosmo

        self.fiducialcosmo = cached_cosmo(self.fiducialparams, self)


_CURRENT: Optional[Config] = None


def init(**kw) -> Config:
    """Drop-in for `cfg.init(...)`; returns (and remembers) the Config."""
    return Config(**kw)


def current() -> Config:
    if _CURRENT is None:
        raise RuntimeError("configuration not initialised: construct a FisherMatrix (or call config.init) first")
    return _CURRENT


def __getattr__(name):  # module-global style access: config.settings, config.specs, ...
    if _CURRENT is not None and hasattr(_CURRENT, name):
        return getattr(_CURRENT, name)
    raise AttributeError(name)

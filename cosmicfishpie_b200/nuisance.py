"""`Nuisance` -- the reference's single nuisance-model class (cosmology/nuisance.py:16-366) as a facade over the
photometric and spectroscopic halves this package keeps next to their jobs (`photo.PhotoNuisance`,
`spectro.SpectroNuisance`).  Same constructor and method names for the GCsp / WL / GCph members; the intensity-mapping
members are out of scope."""
from __future__ import annotations

from . import config as config_module
from .photo import PhotoNuisance
from .spectro import SpectroNuisance


class Nuisance:
    def __init__(self, configuration=None, spectrobiasparams=None, spectrononlinearpars=None, IMbiasparams=None):
        cfg = configuration if configuration is not None else config_module.current()
        self.config, self.specs, self.settings = cfg, cfg.specs, cfg.settings
        self.observables = cfg.obs
        self.specsdir, self.extdir = self.settings.get("specs_dir"), self.settings.get("external_data_dir")
        self.surveyname = self.specs.get("survey_name")
        self.Spectrobiasparams = cfg.Spectrobiasparams if spectrobiasparams is None else spectrobiasparams
        self.Spectrononlinearparams = (getattr(cfg, "Spectrononlinearparams", {}) if spectrononlinearpars is None
                                       else spectrononlinearpars)
        if IMbiasparams is not None and "IM" in self.observables:
            raise NotImplementedError("intensity mapping is out of scope")
        if "GCph" in self.observables or "WL" in self.observables:
            self._photo = PhotoNuisance(cfg)
            self.z = self._photo.z
            self.z_ph = self.z
            if "WL" in self.observables:
                self.lumratio = self._photo.lumratio
        if "GCsp" in self.observables:
            self._spectro = SpectroNuisance(self.specs, self.Spectrobiasparams, self.Spectrononlinearparams)
            self.sp_zbins, self.sp_zbins_inds = self._spectro.sp_zbins, self._spectro.sp_zbins_inds
            self.sp_zbins_mids, self.sp_dndz = self._spectro.sp_zbins_mids, self._spectro.sp_dndz
            self.sp_bias_model = self._spectro.sp_bias_model

    # photometric (nuisance.py:98-230)
    def gcph_bias(self, biaspars, ibin=1):
        return self._photo.gcph_bias(biaspars, ibin)

    def IA(self, IApars, cosmo):
        return self._photo.IA(IApars, cosmo)

    def luminosity_ratio(self):
        return self._photo.lumratio

    # spectroscopic (nuisance.py:232-366)
    def gcsp_zbins(self):
        return self._spectro.gcsp_zbins()

    def gcsp_zbins_mids(self):
        return self._spectro.gcsp_zbins_mids()

    def gcsp_dndz(self):
        return self._spectro.gcsp_dndz()

    def gcsp_bias_at_zm(self):
        return self._spectro.gcsp_bias_at_zm()

    def gcsp_zvalue_to_zindex(self, z):
        return self._spectro.gcsp_zvalue_to_zindex(z)

    def gcsp_bias_at_z(self, z):
        return self._spectro.gcsp_bias_at_z(z)

    def vectorized_gcsp_bias_at_z(self, z):
        return self._spectro.vectorized_gcsp_bias_at_z(z)

    def gcsp_bias_interp(self):
        return self._spectro.gcsp_bias_interp()

    def gcsp_bias_at_zi(self, zi):
        return self._spectro.gcsp_bias_at_zi(zi)

    def gcsp_bias_kscale(self, k, z=None):
        return self._spectro.gcsp_bias_kscale(k, z)

    def vectorized_gcsp_rescale_sigmapv_at_z(self, z, sigma_key="sigmap"):
        return self._spectro.vectorized_gcsp_rescale_sigmapv_at_z(z, sigma_key=sigma_key)

    def gcsp_rescale_sigmapv_at_z(self, z, sigma_key="sigmap"):
        return self._spectro.gcsp_rescale_sigmapv_at_z(z, sigma_key=sigma_key)

    def extra_Pshot_noise(self):
        return self._spectro.extra_Pshot_noise()

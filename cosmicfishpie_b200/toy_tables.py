"""Synthetic Boltzmann-like table bank in the reference's ``external`` layout.

There is no Boltzmann solver on the GPU box (no CAMB/CLASS), and the parity
boundary of this project is *downstream* of the solver: both the reference and
this package consume the same cached tables.  This module
generates a deterministic analytic stand-in with the right shapes and enough
structure to exercise every branch of the hot path:

* flat w0waCDM background, H(z)/c in 1/Mpc
* growth rate f = Omega_m(z)^0.55 with a mild scale dependence (so the k-spline
  of f(z,k) is not trivially constant), growth D from integrating f
* BBKS-like transfer function with damped BAO wiggles (so the Savitzky-Golay
  no-wiggle filter has something to remove), sigma8-normalised
* a HALOFIT-flavoured smooth boost for the non-linear spectrum

Layout written by :func:`write_bank` (one folder per stencil point), as read by
the reference at ``cosmicfishpie/cosmology/cosmology.py:1249-1373`` and selected
by ``:1375-1421``::

    <root>/fiducial_eps_0/{z_values_list,k_values_list,background_Hz,sigma8-z,
                           D_Growth-zk,f_GrowthRate-zk,Plin-zk,Pnonlin-zk}.txt
    <root>/<par>_{pl,mn}_eps_<1p0E-02>/...

All numbers are FP64; nothing here is on the GPU path.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

C_KMS = 299792.458

COSMO_KEYS = ("Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa")

FIDUCIAL = {
    "Omegam": 0.32,
    "Omegab": 0.05,
    "h": 0.67,
    "ns": 0.96,
    "sigma8": 0.815584,
    "w0": -1.0,
    "wa": 0.0,
}

FILE_STEMS = {
    "z": "z_values_list",
    "k": "k_values_list",
    "H": "background_Hz",
    "s8": "sigma8-z",
    "D": "D_Growth-zk",
    "f": "f_GrowthRate-zk",
    "Pl": "Plin-zk",
    "Pnl": "Pnonlin-zk",
    # cold dark matter + baryon ("clustering" tracer) tables, read only when a tracer setting asks for them
    # (config.py:289-304, cosmology.py:1240-1245,1352-1368)
    "s8cb": "sigma8cb-z",
    "fcb": "fcb_GrowthRate-zk",
    "Plcb": "Plincb-zk",
    "Pnlcb": "Pnonlincb-zk",
}


def default_grids(nz: int = 100, nk: int = 500):
    """(z, k) grids of the toy bank: z in [0,3], k in [1e-4, 50] 1/Mpc."""
    return np.linspace(0.0, 3.0, nz), np.logspace(-4.0, np.log10(50.0), nk)


def _E(z, Om, w0, wa):
    a = 1.0 / (1.0 + z)
    de = (1.0 - Om) * a ** (-3.0 * (1.0 + w0 + wa)) * np.exp(-3.0 * wa * (1.0 - a))
    return np.sqrt(Om * (1.0 + z) ** 3 + de)


def _shape(k, Om, Ob, h, ns):
    gam = Om * h * np.exp(-Ob * (1.0 + np.sqrt(2.0 * h) / Om))
    q = (k / h) / gam
    T = (
        np.log(1.0 + 2.34 * q)
        / (2.34 * q)
        * (1.0 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    )
    rs = 147.0 * (0.1426 / (Om * h * h)) ** 0.25 * (0.0224 / (Ob * h * h)) ** 0.13
    wig = 1.0 + 0.05 * (Ob / 0.05) * np.sin(k * rs) * np.exp(-((k / 0.25) ** 1.4))
    return k**ns * T**2 * wig


def make_tables(pars: Dict[str, float], zg: np.ndarray, kg: np.ndarray) -> Dict[str, np.ndarray]:
    """Tables for one cosmology on (zg, kg); keys as in :data:`FILE_STEMS`."""
    Om, Ob, h, ns, s8, w0, wa = (float(pars[k]) for k in COSMO_KEYS)
    Ez = _E(zg, Om, w0, wa)
    H = 100.0 * h * Ez / C_KMS
    # growth: dlnD/dlna = f, integrate on a fine grid
    zf = np.linspace(0.0, float(zg.max()), 4001)
    ff = (Om * (1.0 + zf) ** 3 / _E(zf, Om, w0, wa) ** 2) ** 0.55
    integrand = ff / (1.0 + zf)
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(zf))])
    Dz = np.interp(zg, zf, np.exp(-cum))
    fz = (Om * (1.0 + zg) ** 3 / Ez**2) ** 0.55
    # mild scale dependence (massive-neutrino flavoured), same factor in D and f
    supp = 1.0 - 0.006 * kg / (kg + 0.05)
    fzk = fz[:, None] * (1.0 - 0.004 * kg / (kg + 0.05))[None, :]
    Dzk = Dz[:, None] * supp[None, :]
    # sigma8 normalisation at z=0
    kk = np.logspace(-4.0, 1.5, 4000)
    R = 8.0 / h
    x = kk * R
    W = 3.0 * (np.sin(x) - x * np.cos(x)) / x**3
    yy = kk**3 * _shape(kk, Om, Ob, h, ns) * W**2 / (2.0 * np.pi**2)
    lnk = np.log(kk)
    sig2 = float(np.sum(0.5 * (yy[1:] + yy[:-1]) * np.diff(lnk)))
    P0 = _shape(kg, Om, Ob, h, ns) * s8**2 / sig2
    Pl = Dzk**2 * P0[None, :]
    D2 = kg[None, :] ** 3 * Pl / (2.0 * np.pi**2)
    Pnl = Pl * (1.0 + D2**1.2 / (1.0 + 0.3 * D2**0.6))
    # "clustering" (cb) variants: a smooth scale-dependent boost of the spectra and of the growth rate, as massive
    # neutrinos would leave between total matter and cdm+baryons
    boost = (1.0 + 0.015 * kg / (kg + 0.1))[None, :]
    return {"z": zg, "k": kg, "H": H, "s8": s8 * Dz, "D": Dzk, "f": fzk, "Pl": Pl, "Pnl": Pnl,
            "s8cb": 1.004 * s8 * Dz, "fcb": fzk * (1.0 + 0.003 * kg / (kg + 0.05))[None, :], "Plcb": Pl * boost**2,
            "Pnlcb": Pnl * boost**2}


def eps_string(eps: float) -> str:
    """Folder token for a relative step, e.g. 0.01 -> '1p0E-02' (cosmology.py:1409-1410)."""
    return "{:.1E}".format(abs(eps)).replace(".", "p")


def stencil_points(
    fid: Dict[str, float],
    eps_values: Sequence[float] = (0.01,),
    pars: Optional[Iterable[str]] = None,
):
    """Yield (folder_name, parameter_dict) for the fiducial and every +-eps point.

    Step is relative (fid*eps), or absolute (eps) when the fiducial is 0
    (derivatives.py:140-143, cosmology.py:1393-1396).
    """
    yield "fiducial_eps_0", dict(fid)
    for par in pars if pars is not None else COSMO_KEYS:
        for eps in eps_values:
            for sgn, tag in ((+1.0, "pl"), (-1.0, "mn")):
                p = dict(fid)
                step = fid[par] * eps if fid[par] != 0.0 else eps
                p[par] = fid[par] + sgn * step
                yield f"{par}_{tag}_eps_{eps_string(eps)}", p


def make_bank(
    fid: Optional[Dict[str, float]] = None,
    eps_values: Sequence[float] = (0.01,),
    pars: Optional[Iterable[str]] = None,
    nz: int = 100,
    nk: int = 500,
) -> Dict[str, Dict[str, np.ndarray]]:
    """In-memory bank {folder_name: tables} (what :func:`write_bank` puts on disk)."""
    fid = dict(FIDUCIAL if fid is None else fid)
    zg, kg = default_grids(nz, nk)
    return {name: make_tables(p, zg, kg) for name, p in stencil_points(fid, eps_values, pars)}


def write_bank(root: str, bank: Optional[Dict[str, Dict[str, np.ndarray]]] = None, **kw) -> str:
    """Write a bank as text files in the reference's external layout; returns root."""
    if bank is None:
        bank = make_bank(**kw)
    for name, t in bank.items():
        d = os.path.join(root, name)
        os.makedirs(d, exist_ok=True)
        for key, stem in FILE_STEMS.items():
            np.savetxt(os.path.join(d, stem + ".txt"), t[key])
    return root


def extfiles_for(root: str, eps_values: Sequence[float] = (0.01,), pars=COSMO_KEYS) -> dict:
    """The ``extfiles`` dict both the reference and this package accept (config.py:286-310)."""
    return {
        "directory": root if root.endswith("/") else root + "/",
        "paramnames": list(pars),
        "folder_paramnames": list(pars),
        "k-units": "1/Mpc",
        "r-units": "Mpc",
        "eps_values": list(eps_values),
    }


if __name__ == "__main__":  # pragma: no cover
    import sys

    out = sys.argv[1] if len(sys.argv) > 1 else "toy_bank"
    eps = [float(x) for x in sys.argv[2:]] or [0.01]
    write_bank(out, eps_values=eps)
    print(len(os.listdir(out)), "folders written to", out)

"""CPU: the algebra behind the moment path of the spectroscopic likelihoods (spectro_wedge_moments_kernel,
spectro_legendre_moments_kernel, spectro_moments_apply_kernel), restated in numpy on oracle lattices.

Members of a group differ only in the galaxy bias b and the additive noise m.  P_t(b, m) = (b + F)^2 W + m is a
polynomial in them, both chi^2 are quadratic in P_t, hence chi2_j = v^T M v with v = (b_j^2 - b_0^2, b_j - b_0,
m_j - m_0, 1) and M a sum over the lattice that does not depend on the member.  The GPU test
tests/test_like_gpu.py::test_spectro_moment_path_equals_direct_kernels checks the kernels; this one checks the
identity itself against the direct evaluation, including the exact zero for a member equal to the reference."""
import os

import numpy as np
import pytest
import yaml

from conftest import REPO

SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")


@pytest.fixture(scope="module")
def lattices(toy_bank):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.cosmo import Bank
    from oracle.spectro import DEFAULT_SETTINGS, GalSpectro

    ob = Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    sp = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    fid = ob.cosmo(dict(tt.FIDUCIAL))
    obs = GalSpectro(fid, fid, sp, dict(DEFAULT_SETTINGS), dict(sp["sp_bias_parametrization"]["linear_log"]), {})
    k, mu = np.linspace(0.002, 0.25, 41), np.linspace(0.0, 1.0, 9)
    kk, mm = np.meshgrid(k, mu)
    z = 1.2
    P = {b: obs.observed_Pgg(z, kk, mm, b_i=float(b)) for b in (-1, 0, 1)}
    W = 0.5 * (P[1] + P[-1]) - P[0]          # P(b) = b^2 W + b (2 F W) + F^2 W
    FW2 = 0.5 * (P[1] - P[-1])
    return obs, z, kk, mm, W, FW2, P[0]


def theory(lat, b, m):
    _, _, _, _, W, FW2, P0 = lat
    return b * b * W + b * FW2 + P0 + m


def test_quadratic_form_of_the_bias_is_exact(lattices):
    obs, z, kk, mm = lattices[:4]
    for b in (0.7, 1.46, 2.3):
        assert np.max(np.abs(theory(lattices, b, 0.0) / obs.observed_Pgg(z, kk, mm, b_i=b) - 1)) < 1e-13


def test_wedge_chi2_through_group_moments(lattices):
    rng = np.random.default_rng(5)
    _, _, kk, _, W, FW2, _ = lattices
    w = rng.uniform(0.5, 1.5, size=kk.shape) * kk**2
    b0, m0 = 1.46, 3.0e-4 * 1.0e4
    Pd = theory(lattices, b0, m0)  # the data are the reference member's own spectrum
    g = np.stack([-W / Pd, -FW2 / Pd, -1.0 / Pd, (Pd - theory(lattices, b0, m0)) / Pd])
    M = np.einsum("icd,jcd,cd->ij", g, g, w)
    for _ in range(20):
        b, m = b0 * (1 + 0.05 * rng.standard_normal()), m0 + 50.0 * rng.uniform()
        direct = np.sum(w * ((Pd - theory(lattices, b, m)) / Pd) ** 2)
        v = np.array([b * b - b0 * b0, b - b0, m - m0, 1.0])
        assert abs(v @ M @ v - direct) <= 1e-11 * direct
    assert np.array([0.0, 0.0, 0.0, 1.0]) @ M @ np.array([0.0, 0.0, 0.0, 1.0]) == 0.0  # the reference member: exactly 0
    # data from another member: the reference's residual is the fourth basis function
    Pd = theory(lattices, 1.5, 5.0)
    g = np.stack([-W / Pd, -FW2 / Pd, -1.0 / Pd, (Pd - theory(lattices, b0, m0)) / Pd])
    M = np.einsum("icd,jcd,cd->ij", g, g, w)
    b, m = 1.41, 11.0
    v = np.array([b * b - b0 * b0, b - b0, m - m0, 1.0])
    direct = np.sum(w * ((Pd - theory(lattices, b, m)) / Pd) ** 2)
    assert abs(v @ M @ v - direct) <= 1e-11 * direct


def test_multipole_chi2_through_group_moments(lattices):
    from scipy.special import eval_legendre

    rng = np.random.default_rng(6)
    _, _, kk, mm, W, FW2, _ = lattices
    mu = mm[:, 0]
    lw = np.array([(2 * l + 1) * eval_legendre(l, mu) for l in (0, 2, 4)]) * np.gradient(mu)  # [3][Nmu] quadrature rows
    nk = kk.shape[1]
    A = rng.standard_normal((nk, 3, 3))
    icov = np.einsum("kab,kcb->kac", A, A) + 3.0 * np.eye(3)
    b0, m0 = 1.46, 3.0
    pl = lambda b, m: np.einsum("lr,rk->kl", lw, theory(lattices, b, m))  # noqa: E731  [Nk][3]
    data = pl(1.5, 5.0)
    G = np.stack([-np.einsum("lr,rk->kl", lw, W), -np.einsum("lr,rk->kl", lw, FW2),
                  -np.tile(lw.sum(axis=1), (nk, 1)), data - pl(b0, m0)])          # [4][Nk][3]
    M = np.einsum("ika,kab,jkb->ij", G, icov, G)
    for _ in range(20):
        b, m = b0 * (1 + 0.05 * rng.standard_normal()), m0 + 10.0 * rng.uniform()
        d = data - pl(b, m)
        direct = np.einsum("ka,kab,kb->", d, icov, d)
        v = np.array([b * b - b0 * b0, b - b0, m - m0, 1.0])
        assert abs(v @ M @ v - direct) <= 1e-10 * direct

"""GPU tests added in round 2: device math functions swept over their domains, photo-z parameters in the batched
likelihood, the pipelined batch entry, options["device"], and the NCCL paths on two GPUs (skipped on one)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, base_options

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]


def ulp_err(y, ref):
    return np.max(np.abs(y - ref) / (np.abs(ref) * np.finfo(float).eps))


def test_device_math_functions_match_libm():
    """tab_exp_neg / pair_exp_neg on [-700, 0], tab_log / fast_rcp / fast_rsqrt / fast_div on positive normal numbers
    (csrc/cfp_spectro.cu: their stated domains), against numpy's libm: a few ulp."""
    from cosmicfishpie_b200 import _lib

    ctx = _lib.default_context()
    rng = np.random.default_rng(0)
    x = np.concatenate([-np.logspace(-300, np.log10(700.0), 200000), -rng.uniform(0, 40, 200000), [0.0, -1e-320, -700.0]])
    x = np.maximum(x, -700.0)  # (the functions clamp below -700: exp(-700) ~ 1e-304 stands in for anything smaller)
    for which in (0, 5):
        y = ctx.math_selftest(which, x)
        assert ulp_err(y, np.exp(x)) < 3.0, which  # measured on B200: 1.96 ulp
    assert np.all(ctx.math_selftest(5, np.array([-1e4, -1e300])) < 1e-300)  # clamped, not garbage
    assert np.all(ctx.math_selftest(0, np.array([-1e4, -1e300])) < 1e-300)
    pos = np.concatenate([10.0 ** rng.uniform(-300, 300, 400000), 1.0 + rng.uniform(-1e-3, 1e-3, 1000),
                          [np.finfo(float).tiny, 1.0, 2.0, 0.5, 1e308]])
    y = ctx.math_selftest(1, pos)
    ref = np.log(pos)
    big = np.abs(ref) > 0.5
    assert np.max(np.abs(y[big] - ref[big]) / np.abs(ref[big])) < 4 * np.finfo(float).eps
    assert np.max(np.abs(y[~big] - ref[~big])) < 4e-16  # near 1 the absolute error counts (measured: 7e-17)
    mid = 10.0 ** rng.uniform(-150, 150, 400000)  # 1/x and 1/sqrt(x) stay normal
    assert ulp_err(ctx.math_selftest(2, mid), 1.0 / mid) < 2.0
    assert ulp_err(ctx.math_selftest(3, mid), 1.0 / np.sqrt(mid)) < 3.0
    assert ulp_err(ctx.math_selftest(4, mid), 1.0 / mid) < 2.0


@pytest.fixture(scope="module")
def photo_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    o = base_options(tmp_path_factory.mktemp("r2p"), survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                     survey_name_spectro=False, ell_sampling=24)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=o,
                      observables=["WL", "GCph"], extfiles=extfiles)
    fm.compute()
    return fm


def test_loglike_batch_honours_photoz_parameters(photo_fm):
    """photo-z parameters must not be dropped by the array path (photo_like.py compute_theory updates them)."""
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood

    like = PhotometricLikelihood(cosmo_data=photo_fm)
    keys = ["b2", "AIA", "sigma_b", "zb"]
    fid = np.array([photo_fm.allparams[k] for k in keys])
    mat = np.tile(fid, (4, 1))
    mat[1, 2] *= 1.05
    mat[2, 3] += 0.004
    mat[3] = fid * np.array([1.02, 0.97, 1.0, 1.0]) + np.array([0, 0, 0.002, 0.001])
    batch = like.loglike_batch(mat, keys=keys)
    single = np.array([like.loglike(param_dict=dict(zip(keys, r))) for r in mat])
    assert abs(batch[0]) < 1e-7 and np.all(batch[1:] < -1e-3)  # the photo-z shifts are felt
    assert np.max(np.abs(batch[1:] - single[1:]) / np.abs(single[1:])) < 1e-10


def test_compute_many_equals_compute(extfiles, tmp_path):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    def make(obs, i):
        if obs == "GCsp":
            o = base_options(tmp_path / f"s{i}", spectro_Pk_k_samples=257, spectro_Pk_mu_samples=17)
        else:
            o = base_options(tmp_path / f"p{i}", survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                             survey_name_spectro=False, ell_sampling=16)
        return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "w0": 0.01}, options=o,
                            observables=["GCsp"] if obs == "GCsp" else ["WL", "GCph"], extfiles=extfiles)

    order = ["GCsp", "photo", "photo", "GCsp"]
    many = FisherMatrix.compute_many([make(o, i) for i, o in enumerate(order)])
    for i, o in enumerate(order):
        one = make(o, 10 + i).compute()
        assert list(one.get_param_names()) == list(many[i].get_param_names())
        assert np.array_equal(one.fisher_matrix, many[i].fisher_matrix)


def test_device_option_reaches_the_plans(extfiles, tmp_path):
    """options['device'] must select the device of banks, plans and likelihood helpers."""
    import torch

    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    dev = torch.cuda.device_count() - 1
    o = base_options(tmp_path, spectro_Pk_k_samples=129, spectro_Pk_mu_samples=9, device=dev)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=o, observables=["GCsp"],
                      extfiles=extfiles)
    fm.compute()
    assert fm.ctx.device == dev and fm._spectro_job.plan().ctx.device == dev
    assert fm._spectro_job.cosmo_set.context().device == dev


TWO_GPU_SCRIPT = r'''
import os, sys, tempfile
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CFP_REPO"])
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from cosmicfishpie_b200 import toy_tables as tt
from cosmicfishpie_b200.fishermatrix import FisherMatrix
from cosmicfishpie_b200 import distributed as cdist
from cosmicfishpie_b200.likelihood import SpectroLikelihood
bank = tt.make_bank(eps_values=(0.01,))
ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
       "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}
tmp = tempfile.mkdtemp(prefix="cfp_2gpu_%d_" % rank)
def make(shard):
    o = {"accuracy": 1, "feedback": 0, "code": "external", "outroot": "t", "results_dir": tmp, "survey_name": "Euclid",
         "cosmo_model": "w0waCDM", "derivatives": "3PT", "survey_name_spectro": "Euclid-Spectroscopic-ISTF-Pessimistic",
         "survey_name_photo": False, "spectro_Pk_k_samples": 1025, "spectro_Pk_mu_samples": 33}
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01, "w0": 0.01}, options=o,
                      observables=["GCsp"], extfiles=ext)
    if not shard:
        cdist.shard(fm, 0, 1)
    return fm
F_sharded = make(True).compute().fisher_matrix      # lattice cells split over the ranks + NCCL all-reduce in place
fm1 = make(False)
F_single = fm1.compute().fisher_matrix              # the whole lattice on this rank
err = float(np.max(np.abs(F_sharded - F_single)) / np.max(np.abs(F_single)))
assert np.array_equal(F_sharded == 0, F_single == 0)
like = SpectroLikelihood(cosmoFM_data=fm1)
keys = ["lnbg_1", "lnbg_2", "Ps_3"]
fid = np.array([fm1.allparams[k] for k in keys])
mat = fid + np.array([0.01, 0.01, 5.0]) * np.random.default_rng(3).standard_normal((37, 3))
a = like.loglike_batch(mat, keys=keys, shard=True)
b = like.loglike_batch(mat, keys=keys)
gerr = float(np.max(np.abs(a - b) / np.abs(b)))
if rank == 0:
    print("RESULT", err, gerr, flush=True)
dist.destroy_process_group()
'''


def test_two_gpu_sharded_fisher_and_gathered_likelihood(tmp_path):
    """The NCCL paths on real hardware: a forecast whose lattice is split over two GPUs and all-reduced in place on
    the plan's device buffer equals the single-GPU result to 1e-12; a sharded likelihood batch gathered over the
    ranks equals the unsharded one."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "two_gpu.py"
    script.write_text(TWO_GPU_SCRIPT)
    env = dict(os.environ, CFP_REPO=REPO)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    # (the slices of a sharded batch form other moment groups than the whole batch: rounding-level differences)
    assert float(line[1]) < 1e-12 and float(line[2]) < 1e-11, line


def test_nautilus_sampler_driver_joint_likelihood(extfiles, tmp_path, monkeypatch):
    """The YAML-style driver of the reference (likelihood/sampler.py:86-261) over the vectorised GPU likelihoods, run
    end to end against the stand-in `nautilus` of tools/_stubs: joint photometric + spectroscopic likelihood, chain and
    metadata files as the reference writes them, every proposed batch evaluated by ONE loglike_batch call."""
    monkeypatch.syspath_prepend(os.path.join(REPO, "tools", "_stubs"))
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.sampler import NautilusSampler, load_chain_metadata

    cfg = {
        "name": "t1", "fiducial": dict(tt.FIDUCIAL), "observables": ["WL", "GCph", "GCsp"], "extfiles": extfiles,
        "options": {"accuracy": 1, "feedback": 0, "code": "external", "survey_name": "Euclid", "cosmo_model": "w0waCDM",
                    "survey_name_photo": "Euclid-Photometric-ISTF-Pessimistic",
                    "survey_name_spectro": "Euclid-Spectroscopic-ISTF-Pessimistic", "ell_sampling": 12,
                    "spectro_Pk_k_samples": 129, "spectro_Pk_mu_samples": 9},
        "priors": {"b3": [1.2, 1.5], "AIA": {"type": "gaussian", "loc": 1.72, "scale": 0.05}, "lnbg_2": [0.40, 0.46],
                   "Ps_1": [0.0, 30.0], "not_a_parameter": [0, 1]},
        "sampler_settings": {"n_live": 16, "n_networks": 1, "n_batch": 16, "pool": 4},
    }
    s = NautilusSampler(cfg, base_dir=str(tmp_path))
    assert s.prior_chosen.keys == ["b3", "AIA", "lnbg_2", "Ps_1"]
    naut = s.run()
    assert naut.vectorized and naut.pool is None and naut.n_calls == 3  # three batches, three vectorised calls
    chain, fid, meta, labels = load_chain_metadata(os.path.join(str(tmp_path), "chains", "chains_t1"))
    data = np.loadtxt(chain)
    assert data.shape == (48, 4 + 2) and open(chain).readline().strip() == "# loglike weights b3 AIA lnbg_2 Ps_1"
    assert set(fid) == {"b3", "AIA", "lnbg_2", "Ps_1"} and labels["b3"] == "$b_{3}$"
    for key in ("name", "start_time", "finish_time", "elapsed_time", "evidence_log_z", "chain_file", "priors",
                "sampler_settings", "observables", "cosmo_model", "survey_name", "code", "cosmo_fiducial_params"):
        assert key in meta, key
    # the chain's log-likelihoods are the joint likelihood: photo + spectro, point by point
    ph, sp = s.likelihood.parts
    keys = s.prior_chosen.keys
    for row in data[:3]:
        d = dict(zip(keys, row[:4]))
        assert abs(ph.loglike(param_dict=d) + sp.loglike(param_dict=d) - row[5]) <= 1e-9 * abs(row[5])
    assert np.isclose(np.sum(data[:, 4]), 1.0)
    # a finished run is not repeated: metadata refresh only
    assert s.run() is None


def test_einsum_seam_accepts_user_arrays(photo_fm):
    """SURVEY 8b seam (4): photo_LSS_fishermatrix_einsum(noisy_cls, covmat, derivs) on the CALLER's arrays
    (cosmicfish.py:643-744) -- here the job's own covariance and derivatives passed as copies, and derivatives from the
    `derivatives` engine driven with a user callable: both must give the forecast's matrix."""
    from cosmicfishpie_b200.derivatives import derivatives

    fm = photo_fm
    lss = fm.photo_LSS
    F_job = fm.fisher_array
    noisy, covmat = lss.compute_covmat()
    derivs = lss.compute_derivs()
    cov_copy = [np.array(c, dtype=float) for c in covmat]
    der_copy = {p: {k: np.array(v) for k, v in d.items()} for p, d in derivs.items()}
    F_user = fm.photo_LSS_fishermatrix_einsum(noisy_cls=dict(noisy), covmat=cov_copy, derivs=der_copy)
    assert np.max(np.abs(F_user - F_job) / np.abs(F_job)) < 1e-9
    # numpy restatement of the reference's einsum assembly on the same arrays
    ells = noisy["ells"]
    cols = list(covmat[0].columns)
    lc, dl = 0.5 * (ells[1:] + ells[:-1]), np.diff(ells)
    inv = np.linalg.pinv(np.array(cov_copy))
    der = np.array([[[der_copy[p][f"{a}x{b}"] for b in cols] for a in cols] for p in fm.freeparams])
    ref = np.zeros((len(fm.freeparams),) * 2)
    for li in range(len(lc)):
        d = der[:, :, :, li]
        m2 = np.einsum("ij,pjk,kl->pil", inv[li], d, inv[li])
        ref += np.einsum("pij,qji->pq", d, m2) * (lc[li] + 0.5) * dl[li]
    assert np.max(np.abs(F_user - ref) / np.abs(ref)) < 1e-9
    # the derivative engine on a USER callable (host route) and on the package's own getcls (device route)
    free2 = {"Omegam": 0.01, "AIA": fm.freeparams["AIA"]}
    calls = []

    def my_observable(pars):
        calls.append(dict(pars))
        return lss.getcls(pars)

    d_host = derivatives(my_observable, lss.allparsfid, freeparams=free2, observables_type=["WL", "GCph"]).result
    d_dev = derivatives(lss.getcls, lss.allparsfid, freeparams=free2, observables_type=["WL", "GCph"]).result
    assert len(calls) == 4
    for p in free2:
        for key in ("WL 1xWL 1", "WL 3xGCph 5", "GCph 10xGCph 10"):
            ref_d = derivs[p][key]
            assert np.allclose(d_host[p][key], ref_d, rtol=1e-9, atol=1e-22), (p, key)
            assert np.allclose(d_dev[p][key], ref_d, rtol=1e-12, atol=1e-25), (p, key)
        assert np.array_equal(d_host[p]["ells"], ells)


def test_term_accessors_match_the_oracle(extfiles, tmp_path, toy_bank):
    """SURVEY 8b seams: kaiserTerm / FingersOfGod / dewiggled_pdd (spectro_obs.py:584-676, 811-843), lensing_kernel /
    galaxy_kernel / clsintegral (photo_obs.py:514-615, 878-928) and SpectroDerivs.exact_derivs (spectro_cov.py:510-532),
    evaluated by the device, against the oracle's restatement of the same reference functions."""
    import yaml

    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.spectro import DEFAULT_SETTINGS, GalSpectro

    o = base_options(tmp_path, spectro_Pk_k_samples=65, spectro_Pk_mu_samples=9)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=o, observables=["GCsp"], extfiles=extfiles)
    fm.compute()
    cp = dict(tt.FIDUCIAL, h=tt.FIDUCIAL["h"] * 1.01)
    from cosmicfishpie_b200.spectro import ComputeGalSpectro

    obs = ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL), configuration=fm)
    sp = yaml.safe_load(open(os.path.join(REPO, "cosmicfishpie_b200", "data", "specs",
                                          "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    ob = Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ref = GalSpectro(ob.cosmo(cp), ob.cosmo(dict(tt.FIDUCIAL)), sp, dict(DEFAULT_SETTINGS),
                     dict(sp["sp_bias_parametrization"]["linear_log"]), {})
    z = 1.2
    k, mu = fm.Pk_kmesh, fm.Pk_mumesh
    for name, oname in (("kaiserTerm", "kaiser"), ("FingersOfGod", "fog"), ("dewiggled_pdd", "dewiggled")):
        got, want = getattr(obs, name)(z, k, mu), getattr(ref, oname)(z, k, mu)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12, name
    assert np.max(np.abs(obs.spec_err_lattice(z, k, mu) - ref.spec_err_z(z, k, mu))) < 1e-14
    bterm = ref.nuis.bias_at_z(z)
    rsd = obs.kaiserTerm(z, k, mu, just_rsd=True)   # 1 + (f / b) mu^2
    assert np.max(np.abs(rsd - (1 + (ref.kaiser(z, k, mu) - bterm) / bterm))) < 1e-12
    # exact shot-noise derivative: delta_{i,bin} / P_obs of the last evaluated stencil point
    ex = fm.pk_deriv.exact_derivs("Ps_2")
    assert fm.pk_deriv.exact_derivs("h") is None and np.all(ex[0] == 0) and np.all(ex[1] > 0)
    assert np.array_equal(ex[1], fm.derivs_dict["Ps_2"][1])
    # photometric kernels
    po = base_options(tmp_path / "p", survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False,
                      ell_sampling=12)
    fp = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01}, options=po, observables=["WL", "GCph"],
                      extfiles=extfiles)
    fp.compute()
    cl = fp.photo_obs_fid
    Wwl, Wia = cl.lensing_kernel(cl.z, 4)
    Wgc = cl.galaxy_kernel(cl.z, 7)
    assert Wwl.shape == cl.z.shape and np.all(Wwl >= 0) and Wwl[5] > 0 and np.all(Wia <= 0) and np.all(Wgc >= 0)
    # W^GC = n_i(z) H(z) without bias (photo_obs.py:542): check against the window object and the Hubble table
    H = fp.fiducialcosmo.on_grid("Hubble", cl.z)
    assert np.allclose(Wgc, cl.window.norm_ngal_photoz(cl.z, 7, "GCph") * H, rtol=1e-12, atol=1e-300)
    cl.compute_all()
    assert np.array_equal(cl.clsintegral("WL", 2, "GCph", 5), cl.result["WL 2xGCph 5"])


def test_batched_fisher_postprocessing_on_device():
    """sums, marginal bounds, marginalised matrices and figures of merit of thousands of Fisher matrices in
    one launch each, against numpy and against the result container's own CPU methods
    (analysis/fisher_matrix.py:462-548, 705-733; analysis/fisher_operations.py:177-272)."""
    from cosmicfishpie_b200 import fisher_operations as fo
    from cosmicfishpie_b200.fisher_matrix import fisher_matrix

    rng = np.random.default_rng(5)
    B, n = 3000, 23
    A = rng.standard_normal((B, n, 3 * n))
    scale = 10.0 ** rng.uniform(-2, 3, size=(1, n))
    F = np.einsum("bik,bjk->bij", A, A) * scale[:, :, None] * scale[:, None, :]  # SPD, rows of very different size
    names = [f"p{i}" for i in range(n)]
    names[5], names[6] = "w0", "wa"
    sigma = fo.batch_confidence_bounds(F)
    inv = np.linalg.inv(F)
    from cosmicfishpie_b200.fisher_matrix import confidence_coefficient

    ref_sigma = confidence_coefficient(0.6827) * np.sqrt(np.diagonal(inv, axis1=1, axis2=2))
    assert sigma.shape == (B, n) and np.max(np.abs(sigma / ref_sigma - 1)) < 1e-9  # (north_star asks 1e-4 on sigma)
    s2, marg, fom = fo.batch_marginalise(F, names, ["wa", "p2", "w0"])
    ref_marg = np.linalg.inv(inv[:, [6, 2, 5]][:, :, [6, 2, 5]])
    assert np.max(np.abs(marg - ref_marg) / np.abs(ref_marg).max(axis=(1, 2), keepdims=True)) < 1e-9
    assert np.allclose(fom, np.sqrt(np.linalg.det(ref_marg)), rtol=1e-9)
    assert np.allclose(fo.batch_figure_of_merit(F, names), 1 / np.sqrt(np.linalg.det(inv[:, 5:7, 5:7])), rtol=1e-9)
    # the single-matrix interface of the reference's module and the container's CPU methods
    f0 = fisher_matrix(fisher_matrix=F[0], param_names=names)
    assert np.allclose(sigma[0], f0.get_confidence_bounds(), rtol=1e-10)
    m0 = fo.marginalise(f0, ["w0", "wa"])
    assert m0.param_names == ["w0", "wa"] and m0.name.endswith("_marginal")
    assert np.allclose(m0.fisher_matrix, np.linalg.inv(inv[0][5:7, 5:7]), rtol=1e-9)
    over = fo.marginalise_over(f0, ["p0", "p1"])
    assert over.param_names == names[2:] and np.allclose(over.get_confidence_bounds(), f0.get_confidence_bounds()[2:], rtol=1e-8)
    # sums over the union of parameters
    nb = 15
    Bm = np.einsum("bik,bjk->bij", A[:64, :nb], A[:64, :nb])
    nb_names = names[:7] + [f"q{i}" for i in range(nb - 7)]
    la = [fisher_matrix(fisher_matrix=F[b], param_names=names) for b in range(64)]
    lb = [fisher_matrix(fisher_matrix=Bm[b], param_names=nb_names) for b in range(64)]
    sums = fo.batch_add(la, lb)
    for b in (0, 17, 63):
        ref = la[b] + lb[b]
        assert sums[b].param_names == ref.param_names and np.array_equal(sums[b].fisher_matrix, ref.fisher_matrix)
    # a matrix that is not positive definite
    bad = F[:2].copy()
    bad[1] = -bad[1]
    s = fo.batch_confidence_bounds(bad)
    assert np.all(np.isfinite(s[0])) and np.all(np.isnan(s[1]))


def test_emulated_cosmology_likelihood_and_device_combine(extfiles, toy_bank, tmp_path):
    """with the bank's emulator a likelihood batch moves CONTINUOUSLY in cosmology.  (1) parity at the
    bank nodes: the same numbers as without the emulator; (2) off the nodes: against a likelihood fed with the toy
    model's EXACT tables of each point (stated interpolation error); (3) cfp_bank_combine: the bicubic rows combined on
    the device equal the host fit of the combined tables."""
    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator
    from cosmicfishpie_b200.cosmology import cached_cosmo
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from cosmicfishpie_b200.likelihood import SpectroLikelihood

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    zg, kg = toy_bank["fiducial_eps_0"]["z"], toy_bank["fiducial_eps_0"]["k"]

    class Exact:  # stands in for the emulator: the toy model evaluated at the point itself
        is_node = staticmethod(lambda theta: False)
        tables_at = staticmethod(lambda theta: tt.make_tables({k: theta[k] for k in tt.COSMO_KEYS}, zg, kg))

    def make(ext, sub):
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01, "Omegam": 0.01},
                          options=base_options(tmp_path / sub, spectro_Pk_k_samples=257, spectro_Pk_mu_samples=17),
                          observables=["GCsp"], extfiles=ext)
        fm.compute()
        return fm, SpectroLikelihood(cosmoFM_data=fm)

    fm_plain, like_plain = make(extfiles, "a")
    fm_emu, like_emu = make(dict(extfiles, emulator=emu), "b")
    fm_exact, like_exact = make(dict(extfiles, emulator=Exact()), "c")
    keys = ["Omegam", "h", "sigma8", "w0", "lnbg_1", "lnbg_3"]
    fid = np.array([fm_emu.allparams[k] for k in keys])
    # (1) bank nodes
    nodes = np.tile(fid, (5, 1))
    nodes[1, 0] *= 1.01
    nodes[2, 1] *= 0.99
    nodes[3, 2] *= 1.01
    nodes[4, 3] *= 0.99
    nodes[:, 4:] += 0.01
    a, b = like_emu.loglike_batch(nodes, keys=keys), like_plain.loglike_batch(nodes, keys=keys)
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-12
    # (2) between the nodes, several parameters at once
    rng = np.random.default_rng(21)
    pts = np.tile(fid, (12, 1))
    pts[:, :4] *= 1 + 0.008 * rng.uniform(-1, 1, size=(12, 4))
    pts[:, 4:] += 0.01 * rng.standard_normal((12, 2))
    got, want = like_emu.loglike_batch(pts, keys=keys), like_exact.loglike_batch(pts, keys=keys)
    rel = np.abs(got - want) / np.abs(want)
    assert np.all(np.isfinite(got)) and np.max(rel) < 0.05, rel      # stated error of the diagonal expansion
    snapped = like_plain.loglike_batch(np.where(np.arange(6) < 4, np.round(pts / fid / 0.01) * 0.01 * fid, pts), keys=keys)
    assert np.median(rel) < 0.1 * np.median(np.abs(snapped - want) / np.abs(want))  # ... and far better than snapping
    # (3) device combination of bank rows == host fit of the combined tables
    cs = fm_emu._spectro_job.cosmo_set
    base_pars = [dict(tt.FIDUCIAL)] + [p for _, p in list(tt.stencil_points(dict(tt.FIDUCIAL), (0.01,)))[1:]]
    rows = [cached_cosmo(p, fm_emu.config).packed_row(cs.context(), (_lib.TAB_PLIN, _lib.TAB_F)) for p in base_pars]
    base = _lib.Bank.from_rows(cs.context(), rows)
    theta = dict(tt.FIDUCIAL, Omegam=0.3223, h=0.6725, ns=0.9633, wa=0.004)
    W = emu.weight_matrix([theta], [f for f, _ in tt.stencil_points(dict(tt.FIDUCIAL), (0.01,))])
    comb = base.combined(W)
    zq, kq = rng.uniform(0.9, 1.8, 500), 10 ** rng.uniform(-3, -0.7, 500)
    host = cached_cosmo(theta, fm_emu.config)
    for slot, spl in ((_lib.TAB_PLIN, host.Pk_l), (_lib.TAB_F, host.f_growthrate_zk)):
        dev = comb.eval(len(base_pars), slot, zq, kq)
        ref = spl(zq, kq, grid=False)
        assert np.max(np.abs(dev - ref) / np.abs(ref)) < 1e-11, slot
        assert np.array_equal(comb.eval(3, slot, zq, kq), base.eval(3, slot, zq, kq))  # the base rows are carried over
    comb.close()
    base.close()


def test_emulated_batch_combines_bank_rows_on_the_device(extfiles, toy_bank, tmp_path, monkeypatch):
    """A likelihood batch whose every point has its own (emulated) cosmology: the bank of the batch is combined from
    the tabulated rows on the device (cfp_bank_combine_rows, the emulated tables never exist on the host) and gives
    the log-likelihoods of the route that forms every emulated table on the host and fits it (CFP_NO_COMBINE=1)."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    ext = dict(extfiles, emulator=emu)
    rng = np.random.default_rng(5)

    def batch(fm, keys, n):
        fid = np.array([fm.allparams[k] for k in keys])
        pts = np.tile(fid, (n, 1))
        nc = sum(k in tt.COSMO_KEYS for k in keys)
        pts[:, :nc] *= 1 + 0.008 * rng.uniform(-1, 1, size=(n, nc))
        pts[0, :nc] = fid[:nc]  # (a tabulated cosmology among them: one-hot weights)
        pts[:, nc:] *= 1 + 0.01 * rng.standard_normal((n, len(keys) - nc))
        return pts

    def both(like, pts, keys):
        cz.clear_caches()
        fast = like.loglike_batch(pts, keys=keys)
        used_combine = any(hit[2].blob is None and not hit[2].rows and hit[2].n_cosmo == len(hit[1]) and
                           any(getattr(c, "_emu", None) for c in hit[1]) for hit in cz._BANK_CACHE.values())
        assert used_combine
        cz.clear_caches()
        monkeypatch.setenv("CFP_NO_COMBINE", "1")
        slow = like.loglike_batch(pts, keys=keys)
        monkeypatch.delenv("CFP_NO_COMBINE")
        cz.clear_caches()
        return fast, slow

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01, "Omegam": 0.01},
                      options=base_options(tmp_path / "s", spectro_Pk_k_samples=257, spectro_Pk_mu_samples=17),
                      observables=["GCsp"], extfiles=ext)
    fm.compute()
    keys = ["Omegam", "h", "sigma8", "ns", "lnbg_1", "lnbg_3"]
    fast, slow = both(SpectroLikelihood(cosmoFM_data=fm), batch(fm, keys, 24), keys)
    assert np.all(np.isfinite(fast)) and np.max(np.abs(fast - slow) / np.abs(slow)) < 1e-9

    fp = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01, "Omegam": 0.01},
                      options=base_options(tmp_path / "p", survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                                           survey_name_spectro=False, ell_sampling=24),
                      observables=["WL", "GCph"], extfiles=ext)
    fp.compute()
    keys = ["Omegam", "h", "sigma8", "w0", "b1", "b5", "AIA"]
    fast, slow = both(PhotometricLikelihood(cosmo_data=fp), batch(fp, keys, 40), keys)
    assert np.all(np.isfinite(fast)) and np.max(np.abs(fast - slow) / np.abs(slow)) < 1e-9


def test_nautilus_sampler_over_cosmology_with_the_emulator(extfiles, toy_bank, tmp_path, monkeypatch):
    """A chain that moves in COSMOLOGY (Omegam, h, sigma8 free beside nuisance parameters): with the bank's emulator
    every proposed point gets its own cosmology (the reference would call its Boltzmann solver per point or, with
    external tables, snap to the nearest folder); the batches run through the device-combined banks.  The chain's
    log-likelihoods equal the point-by-point joint likelihood."""
    monkeypatch.syspath_prepend(os.path.join(REPO, "tools", "_stubs"))
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator
    from cosmicfishpie_b200.sampler import NautilusSampler, load_chain_metadata

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    f = tt.FIDUCIAL
    cfg = {
        "name": "cosmo", "fiducial": dict(f), "observables": ["WL", "GCph", "GCsp"], "extfiles": dict(extfiles, emulator=emu),
        "options": {"accuracy": 1, "feedback": 0, "code": "external", "survey_name": "Euclid", "cosmo_model": "w0waCDM",
                    "survey_name_photo": "Euclid-Photometric-ISTF-Pessimistic",
                    "survey_name_spectro": "Euclid-Spectroscopic-ISTF-Pessimistic", "ell_sampling": 12,
                    "spectro_Pk_k_samples": 129, "spectro_Pk_mu_samples": 9},
        "priors": {"Omegam": [f["Omegam"] * 0.993, f["Omegam"] * 1.007], "h": [f["h"] * 0.993, f["h"] * 1.007],
                   "sigma8": [f["sigma8"] * 0.993, f["sigma8"] * 1.007], "b3": [1.2, 1.5], "lnbg_2": [0.40, 0.46]},
        "sampler_settings": {"n_live": 16, "n_networks": 1, "n_batch": 16},
    }
    s = NautilusSampler(cfg, base_dir=str(tmp_path))
    assert s.prior_chosen.keys == ["Omegam", "h", "sigma8", "b3", "lnbg_2"]
    naut = s.run()
    assert naut.vectorized and naut.n_calls == 3
    chain, _, meta, _ = load_chain_metadata(os.path.join(str(tmp_path), "chains", "chains_cosmo"))
    data = np.loadtxt(chain)
    assert data.shape == (48, 5 + 2) and np.all(np.isfinite(data[:, -1]))
    assert np.ptp(data[:, 0]) > 0 and np.ptp(data[:, 1]) > 0  # the chain really moved in cosmology
    ph, sp = s.likelihood.parts
    keys = s.prior_chosen.keys
    for row in data[:3]:
        d = dict(zip(keys, row[:5]))
        assert abs(ph.loglike(param_dict=d) + sp.loglike(param_dict=d) - row[6]) <= 1e-9 * abs(row[6])
    # and the likelihood does depend on the cosmology of the point (it is not snapped to the fiducial folder)
    d0, d1 = dict(zip(keys, data[0, :5])), dict(zip(keys, data[0, :5]))
    d1["Omegam"] = d0["Omegam"] * 1.002
    assert abs(ph.loglike(param_dict=d1) - ph.loglike(param_dict=d0)) > 1e-6 * abs(ph.loglike(param_dict=d0))


def test_emulated_batch_with_tables_in_h_units(extfiles, toy_bank, tmp_path):
    """Tables in h/Mpc: the wavenumber grid of an emulated cosmology scales with its own h, so its bank row cannot be a
    combination of the tabulated rows; the emulated tables are formed (lazily) and fitted on the device instead.  Same
    physics as the 1/Mpc bank: the log-likelihoods of off-node points agree to the emulator's accuracy."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from cosmicfishpie_b200.likelihood import SpectroLikelihood

    # a bank tabulated on ONE grid of k in h/Mpc: every folder samples its own physical wavenumbers k_h * h
    zg = np.asarray(toy_bank["fiducial_eps_0"]["z"])
    k_h = np.asarray(toy_bank["fiducial_eps_0"]["k"]) / tt.FIDUCIAL["h"]
    hbank = {}
    for name, pars in tt.stencil_points(dict(tt.FIDUCIAL), (0.01,)):
        u = dict(tt.make_tables(pars, zg, k_h * pars["h"]))
        u["k"] = k_h
        for key in ("Pl", "Pnl", "Plcb", "Pnlcb"):
            if key in u:
                u[key] = np.asarray(u[key]) * pars["h"] ** 3
        hbank[name] = u
    rng = np.random.default_rng(8)
    keys = ["Omegam", "sigma8", "ns", "lnbg_1"]

    def run(bank, k_units, sub):
        cz.clear_caches()
        ext = dict(extfiles, tables=bank, emulator=TaylorEmulator(bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01))
        ext["k-units"] = k_units
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01},
                          options=base_options(tmp_path / sub, spectro_Pk_k_samples=129, spectro_Pk_mu_samples=9),
                          observables=["GCsp"], extfiles=ext)
        fm.compute()
        like = SpectroLikelihood(cosmoFM_data=fm)
        fid = np.array([fm.allparams[k] for k in keys])
        pts = np.tile(fid, (10, 1))
        pts[:, :3] *= 1 + 0.006 * np.random.default_rng(8).uniform(-1, 1, size=(10, 3))
        pts[:, 3] += 0.01
        ll = like.loglike_batch(pts, keys=keys)
        emulated = [c for c in like.cosmo_theory.pk_obs_fid.cosmo_set.cosmos if c.folder == "<emulated>"]
        return ll, emulated

    ll_h, emu_h = run(hbank, "h/Mpc", "h")
    assert emu_h and all(c._emu is None for c in emu_h)      # not combinable: formed and fitted per cosmology
    ll_m, emu_m = run(toy_bank, "1/Mpc", "m")
    assert emu_m and all(c._emu is not None for c in emu_m)  # combined on the device
    cz.clear_caches()
    # (the two banks sample the spectra at different wavenumbers and expand different table values in h)
    assert np.all(np.isfinite(ll_h)) and np.max(np.abs(ll_h - ll_m) / np.abs(ll_m)) < 2e-2

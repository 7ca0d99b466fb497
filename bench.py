#!/usr/bin/env python
"""Benchmark of the observable -> Fisher hot path (BASELINE.json metric: Fisher matrices / s).

    python bench.py --gpus 1 --steps K --warmup W               # this package on N GPUs
    python bench.py --impl reference --steps K --warmup W       # CPU arm (oracle port of the reference)

One STEP = one Euclid forecast: a 3x2pt photometric Fisher (WL+GCph+XC, 13+13 tomographic bins,
lmax_WL = 5000, w0waCDM 7 cosmological + 13 galaxy-bias + 3 IA parameters, 100 multipoles x 200 z) PLUS a
GCsp Fisher (4 z-bins, 7 cosmological + 4 bias + 4 shot-noise parameters, 3PT stencil, fine lattice
N_mu x N_k = 129 x 8193), summed into one combined Fisher matrix.  Inputs are synthetic (deterministic toy
Boltzmann tables, cosmicfishpie_b200/toy_tables.py) in the reference's `external` table format.

Numbers in the JSON line
  value      combined Fisher matrices / s with every input resident in HBM (plans uploaded, device timing by
             CUDA events on the launch stream, L2 flushed between steps); N GPUs: N independent forecasts per
             step (weak scaling), `strong_scaling` = one forecast sharded over the GPUs + NCCL all-reduce.  The
             two probes of a forecast run on two streams; the timed steps' results are checked against the API run
  e2e        the same metric through the public API `FisherMatrix(...).compute()` from HOST tables: per-job host
             preparation, pinned H2D of the table bank and job arrays, kernels, D2H of the Fisher matrices and
             file export inside the timed region; the spline facades of the cached tables are reused between
             steps (`e2e_cold` re-fits them every step, which is what the reference does per Fisher matrix)
  likelihood batched log-likelihood evaluations / s through `loglike_batch` (photometric 3x2pt, spectroscopic wedges)
  cabi_e2e   through the C ABI only (cfp_*_plan_create from host buffers + run + fetch)
  roofline   dominant kernel (spectro_fisher_kernel) against the measured FP64 DFMA peak of this GPU; the
             HBM view of the API-preserving (materialising) mode is in `roofline_hbm`
  cpu_baseline  the oracle (numpy/scipy port of the reference algorithm) on one host core, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402

P7 = list(tt.COSMO_KEYS)
PHOTO_SPEC = "Euclid-Photometric-13bins-lmax5000"
SPECTRO_SPEC = "Euclid-Spectroscopic-ISTF-Pessimistic"
NMU, NK = 129, 8193
WORKLOAD = f"euclid_3x2pt_13bins_lmax5000_npar23 + gcsp_4bins_npar15_3PT_lattice{NMU}x{NK}"
# FP64 operations spectro_fisher_kernel EXECUTES per lattice cell and bin, frozen from the ncu source-page counters of
# the final kernel (predicated-on thread instructions, DADD + DMUL + 2 DFMA; profiles/r1e_*): a full evaluation of
# P_obs of one stencil point incl. its log (F_FULL), a point that differs from the fiducial only in the bias and
# re-uses its look-ups (F_LIGHT), and the per-cell epilogue (derivative scaling, V_eff, weights: F_FIXED).
# 15 full + 2 light + fixed = 1510 measured on the bench workload (DFMA 2.551e9, DMUL 0.903e9, DADD 0.379e9 thread
# instructions per launch).  The Gram update adds 128 flop per 8x8 tile (DMMA).
F_FULL, F_LIGHT, F_FIXED = 93.35, 30.0, 50.0
# dram__bytes_read.sum + dram__bytes_write.sum of one spectro_fisher_kernel launch (ncu --set full, profiles/r1e)
TRAFFIC_BYTES = 3.24e6  # 2.99 MB read + 0.25 MB written


def ext_dict(bank):
    return {"tables": bank, "directory": "<memory>", "paramnames": P7, "folder_paramnames": P7,
            "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}


def base_options(tmp, **kw):
    o = {"accuracy": 1, "feedback": 0, "code": "external", "outroot": "bench", "results_dir": tmp,
         "survey_name": "Euclid", "cosmo_model": "w0waCDM", "derivatives": "3PT"}
    o.update(kw)
    return o


def make_photo_fm(bank, tmp):
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                        options=base_options(tmp, survey_name_photo=PHOTO_SPEC, survey_name_spectro=False),
                        observables=["WL", "GCph"], extfiles=ext_dict(bank))


def make_spectro_fm(bank, tmp, nmu=NMU, nk=NK):
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                        options=base_options(tmp, survey_name_spectro=SPECTRO_SPEC, survey_name_photo=False,
                                             spectro_Pk_k_samples=nk, spectro_Pk_mu_samples=nmu),
                        observables=["GCsp"], extfiles=ext_dict(bank))


# ------------------------------------------------------------------------------ CPU arm (oracle)
def oracle_sample(bank, nmu_s=17, nk_s=1025):
    """One bounded CPU sample of the step: the full 3x2pt Fisher plus the GCsp Fisher on a 17x1025 lattice,
    whose time is scaled by the cell ratio to the 129x8193 workload lattice.  Returns seconds per step."""
    import yaml

    from oracle.cosmo import Bank
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher
    from oracle.spectro import spectro_fisher

    specdir = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
    ob = Bank(bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ph = finish_photo_specs(yaml.safe_load(open(os.path.join(specdir, PHOTO_SPEC + ".yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias, ia = default_photo_bias(ph), dict(ph["IA_params"])
    fp = {p: 0.01 for p in P7}
    fp.update({k: ph["vary_ph_bias"] for k in bias if k != "bias_model"})
    fp.update({k: ph["vary_IA_pars"] for k in ia if k != "IA_model"})
    t0 = time.perf_counter()
    photo_fisher(ob, ph, None, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(ph["photo_z_params"]), ia, bias, fp, lum)
    t_photo = time.perf_counter() - t0
    sp = yaml.safe_load(open(os.path.join(specdir, SPECTRO_SPEC + ".yaml")))["specifications"]
    sb, ps = dict(sp["sp_bias_parametrization"]["linear_log"]), dict(sp["shot_noise_parametrization"]["constant"])
    fs = {p: 0.01 for p in P7}
    fs.update({k: 1e-4 for k in list(sb) + list(ps)})
    t0 = time.perf_counter()
    spectro_fisher(ob, sp, {"spectro_Pk_k_samples": nk_s, "spectro_Pk_mu_samples": nmu_s}, dict(tt.FIDUCIAL), sb, ps,
                   {}, fs)
    t_sp = time.perf_counter() - t0
    scale = (NMU * NK) / float(nmu_s * nk_s)
    return t_photo + t_sp * scale, {"photo_s": t_photo, "spectro_sample_s": t_sp, "spectro_scale": scale}


def _oracle_worker(_):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    bank = tt.make_bank(eps_values=(0.01,))
    return oracle_sample(bank)[0]


def run_reference(args):
    """`--impl reference`: the CPU implementation on all host cores (independent processes, which is how
    the reference parallelises -- Nautilus pool=N; a single Fisher is single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup and 0):  # the CPU arm needs no warm-up: caches are per process
            pool.map(_oracle_worker, range(cores))
        per_step = []
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter()
            times = pool.map(_oracle_worker, range(cores))
            wall = time.perf_counter() - t0
            # every worker processed the bounded sample; its scaled step time is what it returned
            per_step.append((wall, max(times)))
    t_step = float(np.mean([t for _, t in per_step]))
    value = cores / t_step
    sample = ("per process: full 3x2pt 13+13-bin Fisher + GCsp Fisher on a 17x1025 lattice, spectro time scaled x"
              f"{(NMU * NK) / (17 * 1025):.1f} (cell ratio) to the 129x8193 workload lattice; {cores} independent "
              "single-threaded processes")
    line = {"impl": "reference", "metric": "fisher_matrices_per_s", "value": value, "unit": "Fisher/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step / cores, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "Fisher/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Fisher/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


# ------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_gpu(args):
    import torch

    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200 import distributed as cdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix=f"cfp_bench_r{rank}_")
    bank = tt.make_bank(eps_values=(0.01,))
    ctx = _lib.default_context(local)

    # ---- resident plans (host prep + upload outside the timed region of `value`) -------------------
    fm_ph = make_photo_fm(bank, tmp)
    fm_sp = make_spectro_fm(bank, tmp)
    fm_ph.compute()
    fm_sp.compute()
    pplan = fm_ph.photo_LSS._job.plan()
    splan = fm_sp._spectro_job.plan()
    ncell = NMU * NK
    nbins = pplan.Nl - 1
    cb, ce = cdist.split_range(ncell, rank, world)
    lb, le = cdist.split_range(nbins, rank, world)
    splan.set_slice(cb, ce)
    pplan.set_slice(lb, le)
    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream
    Fsp = cdist.device_view(splan.fisher_device_ptr(), (splan.Z, splan.npar, splan.npar), local)
    Fph = cdist.device_view(pplan.fisher_device_ptr(), (pplan.npar, pplan.npar), local)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    stream2 = torch.cuda.Stream()
    s2ptr = stream2.cuda_stream
    two_streams = os.environ.get("CFP_BENCH_ONE_STREAM", "0") != "1"

    def step(sharded=False):
        # The two probes of a forecast are independent: the photometric plan runs on a second stream, forked from and
        # joined to the launch stream (whose events time the step), so that its small launch-bound kernels overlap the
        # start and the tail of the lattice kernel instead of preceding it.
        if two_streams:
            stream2.wait_stream(stream)
            pplan.run(stream=s2ptr)
            splan.run(materialize=0, stream=sptr)
            stream.wait_stream(stream2)
        else:
            pplan.run(stream=sptr)
            splan.run(materialize=0, stream=sptr)
        if sharded and world > 1:
            with torch.cuda.stream(stream):
                torch.distributed.all_reduce(Fph)
                torch.distributed.all_reduce(Fsp)

    def timed_loop(sharded):
        """W warm-up steps, then exactly K steps timed with CUDA events on the launch stream (L2 flushed between
        steps, outside the event pairs); returns (ms per step = max over ranks, kernels launched)."""
        if sharded:
            splan.set_slice(cb, ce)
            pplan.set_slice(lb, le)
        else:
            splan.set_slice(0, ncell)
            pplan.set_slice(0, nbins)
        for _ in range(max(3, args.warmup)):
            step(sharded)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        n0 = ctx.launches
        with torch.cuda.stream(stream):
            for a, b in ev:
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
                a.record(stream)
                step(sharded)
                b.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()) / args.steps, ctx.launches - n0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # headline: a batch of independent forecasts, one per GPU per step (weak scaling, no collective)
    ms_per_step, launches_timed = timed_loop(sharded=False)
    value = world * 1e3 / ms_per_step
    # what the timed steps computed is the forecast: compare the resident plans' results with the public-API run
    torch.cuda.synchronize()
    step_check = max(
        float(np.max(np.abs(Fph.cpu().numpy() - fm_ph.fisher_array)) / np.max(np.abs(fm_ph.fisher_array))),
        float(np.max(np.abs(Fsp.cpu().numpy().sum(axis=0) - fm_sp.fisher_array)) / np.max(np.abs(fm_sp.fisher_array))))
    if not step_check < 1e-12:
        raise RuntimeError(f"timed steps do not reproduce the API result (max rel diff {step_check:.3e})")
    # one forecast split over the GPUs (lattice cells / multipole bins) + NCCL all-reduce of the partial Fishers
    strong = None
    if world > 1:
        ms_strong, _ = timed_loop(sharded=True)
        strong = {"value": 1e3 / ms_strong, "unit": "Fisher/s", "ms_per_step": ms_strong, "scaling": "strong",
                  "what": "ONE forecast per step split over the ranks (spectro lattice cells, photo multipole bins), "
                          "one NCCL all-reduce of the partial Fisher matrices per probe inside the timed region"}
    splan.set_slice(cb, ce)
    pplan.set_slice(lb, le)

    # ---- dominant kernel alone: spectro_fisher_kernel ------------------------------------------------
    kms = []
    for _ in range(max(3, min(args.steps, 10))):
        with torch.cuda.stream(stream):
            flush.zero_()
        splan.run(materialize=0, stream=sptr)
        kms.append(splan.kernel_ms())
    k_ms = float(np.mean(kms))
    pplan.run(stream=sptr)
    photo_ms = {"limber_table": pplan.kernel_ms(0), "cls_dmma": pplan.kernel_ms(1), "fisher": pplan.kernel_ms(2)}
    # keep the identical step loop running for >= 0.6 s so that nvidia-smi (20 ms period, slow to start) gets
    # samples under this exact load even when the timed region is only a few milliseconds long
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        clk["window"] = "timed region + 0.6 s of the identical step loop"
    job = fm_sp._spectro_job
    from cosmicfishpie_b200._lib import SP_BIAS
    n_full = n_light = 0
    for z in range(splan.Z):
        own = [s_ for s_ in range(1, splan.S) if job.alias[s_, z] == s_]
        other = [q for q in range(job.scal.shape[2]) if q != SP_BIAS]
        light = [s_ for s_ in own if np.array_equal(job.scal[s_, z, other], job.scal[0, z, other])]
        n_light += len(light)
        n_full += 1 + len(own) - len(light)
    ncell = ce - cb
    evals = (n_full + n_light) * ncell
    nb8 = (splan.npar + 7) // 8
    gram = 128.0 * nb8 * (nb8 + 1) / 2
    flops = (F_FULL * n_full + F_LIGHT * n_light + (F_FIXED + gram) * splan.Z) * ncell
    peak_dfma = ctx.peak_fp64(0)
    peak_dmma = ctx.peak_fp64(1)
    achieved = flops / (k_ms * 1e-3) / 1e12
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s", "frac": achieved / peak_dfma,
            "traffic": TRAFFIC_BYTES if world == 1 else None, "kernel": "spectro_fisher_kernel", "kernel_ms": k_ms,
            "peak_source": "measured live: cfp_peak_fp64 DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry)",
            "peak_dmma_tflops": peak_dmma, "flop_per_full_eval": F_FULL, "flop_per_light_eval": F_LIGHT,
            "flop_per_cell_fixed": F_FIXED + gram, "full_evals_per_cell": n_full, "light_evals_per_cell": n_light,
            "cell_evals": evals,
            "what": "FP64 operations the kernel executes (ncu-frozen per-evaluation counts x evaluations) / "
                    "CUDA-event kernel time, against the DFMA rate measured live on this GPU"}
    # HBM view: the API-preserving mode writes derivs[npar][Z][cells] + veff[Z][cells] once
    with torch.cuda.stream(stream):
        flush.zero_()
    splan.run(materialize=_lib.MAT_DERIVS, stream=sptr)
    m_ms = splan.kernel_ms()
    bytes_alg = 8.0 * splan.Z * (ce - cb) * (splan.npar + 1)
    hbm_peak = 6551.7
    try:
        hbm_peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json (measured)"
    except Exception:
        hbm_peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    roof_hbm = {"bound": "hbm", "achieved": bytes_alg / (m_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_alg / (m_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "kernel": "spectro_fisher_kernel "
                "(materialising derivs+veff)", "kernel_ms": m_ms, "peak_source": src}
    splan.set_slice(0, ncell)
    pplan.set_slice(0, nbins)

    if args.profile:
        if world > 1:
            torch.distributed.destroy_process_group()
        return ({"profile_run": True, "ms_per_step": ms_per_step, "kernel_ms": k_ms, "photo_kernel_ms": photo_ms,
                 "roofline": roof} if rank == 0 else None)

    # ---- end to end through the public API, from host tables -----------------------------------------
    from cosmicfishpie_b200 import cosmology as ccos

    def api_step(cold):
        if cold:
            ccos.clear_caches()  # tables -> FITPACK splines are re-fitted, as the reference does per Fisher
        a = make_photo_fm(bank, tmp)
        cdist.shard(a, 0, 1)  # every rank computes its own full forecast (weak scaling, like `value`)
        Fa = a.compute()
        b = make_spectro_fm(bank, tmp)
        cdist.shard(b, 0, 1)
        Fb = b.compute()
        tot = Fa + Fb
        h2d = (a.photo_LSS._job.plan().h2d_bytes + a.photo_LSS._job.cosmo_set.bank().nbytes
               + b._spectro_job.plan().h2d_bytes + b._spectro_job.cosmo_set.bank().nbytes)
        d2h = 8 * (a.fisher_array.size + b.pk_fisher_per_bin.size)
        a.photo_LSS._job.close()
        b._spectro_job.close()
        a.photo_LSS._job.cosmo_set.bank().close()
        b._spectro_job.cosmo_set.bank().close()
        return tot, h2d, d2h

    def time_api(cold, n):
        api_step(cold)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            out = api_step(cold)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        tt_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(tt_, op=torch.distributed.ReduceOp.MAX)
        return float(tt_.item()), out

    n_e2e = max(2, min(args.steps, 5))
    t_cold, _ = time_api(True, max(2, n_e2e // 2))
    t_api, (Ftot, h2d, d2h) = time_api(False, n_e2e)

    # ---- batched likelihoods (BASELINE.json: likelihood evals / s), each rank its own batch (weak) --------
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood

    def like_block(like, keys, allp, B, sigma):
        rng = np.random.default_rng(1234 + rank)
        fid = np.array([allp[k] for k in keys])
        mat = fid * (1.0 + sigma * rng.standard_normal((B, len(keys))))
        like.loglike_batch(mat, keys=keys)  # warm-up at the timed batch size (the memory pool grows once)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll = like.loglike_batch(mat, keys=keys)
        dt = time.perf_counter() - t0
        tt_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(tt_, op=torch.distributed.ReduceOp.MAX)
        return {"batch_per_gpu": B, "evals_per_s": world * B / float(tt_.item()), "device_ms_last_chunk": ctx.last_run_ms,
                "finite": bool(np.all(np.isfinite(ll)))}

    Lp = PhotometricLikelihood(cosmo_data=fm_ph)
    kp = [k for k in fm_ph.freeparams if k not in P7]
    like_photo = like_block(Lp, kp, fm_ph.allparams, 16384, 0.02)
    Ls = SpectroLikelihood(cosmoFM_data=fm_sp)
    ks = [k for k in fm_sp.freeparams if k.startswith("lnbg")]
    like_spectro = like_block(Ls, ks, fm_sp.allparams, 16384, 0.01)
    Ll = SpectroLikelihood(cosmoFM_data=fm_sp, leg_flag="legendre")
    like_legendre = like_block(Ll, ks, fm_sp.allparams, 16384, 0.01)

    # ---- C ABI only: plan_create from host buffers + run + fetch -------------------------------------
    def cabi_step():
        jp, js = fm_ph.photo_LSS._job, fm_sp._spectro_job
        jp.close()
        js.close()
        p1, p2 = jp.plan(), js.plan()
        p1.run(stream=ctx.aux_stream)  # the two probes are independent: second stream of the context
        p2.run()
        p1.fetch()
        p2.fetch()

    cabi_step()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        cabi_step()
    t_cabi = (time.perf_counter() - t0) / n_e2e

    line = None
    if rank == 0:
        cpu = None
        if world == 1 or True:
            t_cpu, detail = oracle_sample(bank)
            cpu = {"value": 1.0 / t_cpu, "unit": "Fisher/s", "cores": 1, "kind": "port",
                   "sample": "full 3x2pt 13+13-bin Fisher (%.1f s) + GCsp Fisher on a 17x1025 lattice (%.2f s) scaled x%.1f "
                             "by cell count to 129x8193" % (detail["photo_s"], detail["spectro_sample_s"], detail["spectro_scale"])}
        line = {
            "metric": "fisher_matrices_per_s", "value": value, "unit": "Fisher/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2_flush_between_steps": True, "streams_per_step": 2 if two_streams else 1,
                       "partition": "a batch of independent forecasts, one per GPU per step, no collective "
                                    "(strong_scaling: one forecast sharded over the GPUs + NCCL all-reduce)"},
            "strong_scaling": strong,
            "clocks": clk,
            "e2e": {"value": world / t_api, "unit": "Fisher/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "public API per step: FisherMatrix(...) + compute() for 3x2pt and for GCsp, summed; per-job "
                            "host prep, H2D of the table bank and job arrays, kernels, D2H, file export. The spline "
                            "facades of the cached Boltzmann tables are reused across steps (north_star: Boltzmann input "
                            "is computed once on the host and cached); e2e_cold re-fits them every step",
                    "s_per_step": t_api},
            "e2e_cold": {"value": world / t_cold, "unit": "Fisher/s", "s_per_step": t_cold,
                         "what": "as e2e but every cache cleared each step (tables -> FITPACK fits -> bank), i.e. what "
                                 "the reference redoes for every Fisher matrix"},
            "cabi_e2e": {"value": world / t_cabi, "unit": "Fisher/s", "s_per_step": t_cabi,
                         "what": "cfp_*_plan_create(host buffers) + run + fetch, table bank resident"},
            "likelihood": {"photo_3x2pt_13bins": like_photo, "spectro_wedges_129x8193": like_spectro,
                           "spectro_legendre_129x8193": like_legendre,
                           "scaling": "weak", "what": "loglike_batch through the public API (host prep + H2D + kernels + "
                                                      "D2H), nuisance parameters drawn around the fiducial"},
            "gpu_launches": int(launches_timed),
            "roofline": roof, "roofline_hbm": roof_hbm, "photo_kernel_ms": photo_ms, "cpu_baseline": cpu,
            "fisher_trace": float(np.trace(Ftot.fisher_matrix)), "timed_step_vs_api_max_rel_diff": step_check,
        }
    if world > 1:
        torch.distributed.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profile", action="store_true",
                    help="resident-plan loop only (no e2e / CPU baseline): the command line used under ncu")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version there):
    # point fd 1 at stderr for the duration of the run and emit the result on the saved descriptor.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_out, 1)
    if line is not None:
        os.write(1, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()

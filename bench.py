#!/usr/bin/env python
"""Benchmark of the observable -> Fisher hot path (BASELINE.json metric: Fisher matrices / s and likelihood evals / s).

    python bench.py --gpus 1 --steps K --warmup W               # this package on N GPUs
    python bench.py --impl reference --steps K --warmup W       # CPU arm: the reference itself on all host cores

One STEP = one Euclid forecast: a 3x2pt photometric Fisher (WL+GCph+XC, 13+13 tomographic bins,
lmax_WL = 5000, w0waCDM 7 cosmological + 13 galaxy-bias + 3 IA parameters, 100 multipoles x 200 z) PLUS a
GCsp Fisher (4 z-bins, 7 cosmological + 4 bias + 4 shot-noise parameters, 3PT stencil, fine lattice
N_mu x N_k = 129 x 8193), summed into one combined Fisher matrix.  Inputs are synthetic (deterministic toy
Boltzmann tables, cosmicfishpie_b200/toy_tables.py) in the reference's `external` table format.

Numbers in the JSON line
  value      combined Fisher matrices / s with every input resident in HBM (plans uploaded, device timing by
             CUDA events on the launch stream, L2 flushed between steps); N GPUs: N independent forecasts per
             step (weak scaling).  The two probes of a forecast run on two streams; the timed steps' results are
             checked against the API run.
  strong_scaling_value   ONE forecast sharded over the GPUs (lattice cells / multipole bins) + one fused NCCL all-reduce
             of [F_photo | F_spectro] inside the timed region; the reduced result is asserted equal to the
             single-GPU forecast (1e-12) before the number is reported.
  e2e        the same metric through the public API from HOST tables: FisherMatrix(...) per probe, per-job host
             preparation, H2D of the job arrays (and of the raw tables when their fitted bank is not yet resident on
             the device), kernels, D2H of the Fisher matrices and file export inside the timed region -- a batch of forecasts through `FisherMatrix.compute_many` (host
             preparation of forecast i+1 overlaps the kernels of forecast i); `e2e_single` = one blocking
             `compute()` per probe, `e2e_cold` = every cache cleared each step (the reference re-fits its splines
             for every Fisher matrix)
  likelihood batched log-likelihood evaluations / s through `loglike_batch` (photometric 3x2pt, spectroscopic wedges
             and multipoles; nuisance-only batches AND batches in which every point has its own cosmology), and
             `sharded`: one batch split over the ranks with an all-gather of the log-likelihoods
  roofline   dominant kernel (spectro_fisher_pairs_kernel) against the measured FP64 DFMA peak of this GPU
  cpu_baseline  the oracle (numpy/scipy port of the reference algorithm) on ONE host core on the full workload
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
SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
NMU, NK = 129, 8193
WORKLOAD = f"euclid_3x2pt_13bins_lmax5000_npar23 + gcsp_4bins_npar15_3PT_lattice{NMU}x{NK}"
# FP64 operations the lattice pass EXECUTES, frozen from the ncu counters of the final kernels (predicated-on thread
# instructions, DADD + DMUL + 2 DFMA; profiles/r2_*): per warp step of spectro_fisher_pairs_kernel (four cells x
# eight parameter slots: 32 lanes x [2 evaluations + the pair's finishing]) plus its one DMMA.8x8x4 (512 flop), and per
# cell of the spectro_weights_kernel pre-pass.  warp steps = Z x ceil(Nk / 4) x rows (+1 per task: pipeline fill).
# Measured on the bench forecast (profiles/r2f_kernels.md): 4.4594e9 vector flop in 1 073 676 warp steps, 4.6034e8 in
# 4 227 588 cells of the pre-pass.
F_PAIR_STEP = 4153.4 + 512.0
F_WEIGHT_CELL = 108.9
# dram__bytes_read.sum + dram__bytes_write.sum of one spectro_fisher_pairs_kernel launch: `ncu --set full` with its
# default cache flush before the kernel, i.e. the 67.7 MB {weight, 1/P} pairs of the pre-pass come from HBM (profiles/
# r2b); in the running step they are still L2-resident and the kernel touches 0.41 MB of HBM (--cache-control none)
TRAFFIC_BYTES = 70.9e6


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


def make_spectro_fm(bank, tmp, nmu=NMU, nk=NK, **kw):
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                        options=base_options(tmp, survey_name_spectro=SPECTRO_SPEC, survey_name_photo=False,
                                             spectro_Pk_k_samples=nk, spectro_Pk_mu_samples=nmu, **kw),
                        observables=["GCsp"], extfiles=ext_dict(bank))


# ------------------------------------------------------------------------------ CPU arm
def _oracle_inputs(bank):
    import yaml

    from oracle.cosmo import Bank
    from oracle.photo import default_photo_bias, finish_photo_specs

    ob = Bank(bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ph = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, PHOTO_SPEC + ".yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    sp = yaml.safe_load(open(os.path.join(SPECDIR, SPECTRO_SPEC + ".yaml")))["specifications"]
    return ob, ph, lum, sp, default_photo_bias(ph), dict(ph["IA_params"])


def oracle_forecast(bank, nmu=NMU, nk=NK, stencil="3PT", photo=True):
    """One forecast of the workload with the oracle port (numpy/scipy restatement of the reference), one core.
    Returns (seconds, detail)."""
    from oracle.photo import photo_fisher
    from oracle.spectro import spectro_fisher

    ob, ph, lum, sp, bias, ia = _oracle_inputs(bank)
    t_photo = 0.0
    if photo:
        fp = {p: 0.01 for p in P7}
        fp.update({k: ph["vary_ph_bias"] for k in bias if k != "bias_model"})
        fp.update({k: ph["vary_IA_pars"] for k in ia if k != "IA_model"})
        t0 = time.perf_counter()
        photo_fisher(ob, ph, None, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(ph["photo_z_params"]), ia, bias, fp, lum)
        t_photo = time.perf_counter() - t0
    sb, ps = dict(sp["sp_bias_parametrization"]["linear_log"]), dict(sp["shot_noise_parametrization"]["constant"])
    fs = {p: 0.01 for p in P7}
    fs.update({k: 1e-4 for k in list(sb) + list(ps)})
    t0 = time.perf_counter()
    spectro_fisher(ob, sp, {"spectro_Pk_k_samples": nk, "spectro_Pk_mu_samples": nmu}, dict(tt.FIDUCIAL), sb, ps, {}, fs,
                   stencil=stencil)
    t_sp = time.perf_counter() - t0
    return t_photo + t_sp, {"photo_s": t_photo, "spectro_s": t_sp}


def oracle_likelihoods(bank):
    """Seconds per single log-likelihood evaluation of the oracle port on one core: photometric 3x2pt 13+13 bins, and
    spectroscopic wedges / multipoles on the 129 x 8193 lattice (a point with its own cosmology)."""
    from oracle.likelihood import PhotoLikelihood, SpectroLikelihoodOracle

    ob, ph, lum, sp, bias, ia = _oracle_inputs(bank)
    L = PhotoLikelihood(ob, ph, None, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(ph["photo_z_params"]), ia, bias, lum)
    pts = [{"b1": bias["b1"] * 1.02, "AIA": ia["AIA"] * 0.97}, {"Omegam": tt.FIDUCIAL["Omegam"] * 1.01, "b5": bias["b5"] * 0.99}]
    t0 = time.perf_counter()
    for p in pts:
        L.loglike(p)
    t_photo = (time.perf_counter() - t0) / len(pts)
    sb, ps = dict(sp["sp_bias_parametrization"]["linear_log"]), dict(sp["shot_noise_parametrization"]["constant"])
    S = SpectroLikelihoodOracle(ob, sp, {"spectro_Pk_k_samples": NK, "spectro_Pk_mu_samples": NMU}, dict(tt.FIDUCIAL), sb, ps, {})
    pld = S.legendre(S.data)
    _, icov = S.legendre_cov(pld)
    t0 = time.perf_counter()
    th = S.theory({"h": tt.FIDUCIAL["h"] * 1.01, "lnbg_2": sb["lnbg_2"] + 0.01})
    S.wedge_chi2(th)
    t_w = time.perf_counter() - t0
    t0 = time.perf_counter()
    S.legendre_chi2(pld, S.legendre(th), icov)
    t_l = t_w + time.perf_counter() - t0
    return {"photo_3x2pt_13bins_evals_per_s": 1.0 / t_photo, "spectro_wedges_129x8193_evals_per_s": 1.0 / t_w,
            "spectro_legendre_129x8193_evals_per_s": 1.0 / t_l, "cores": 1}


def _reference_importable():
    ref = os.path.join(REPO, "oracle", "_ref")
    return os.path.isdir(os.path.join(ref, "cosmicfishpie"))


def _reference_worker(args):
    """One full forecast of the workload with the UNMODIFIED reference (oracle/_ref, installed by
    oracle/build_ref.sh) if it travelled with the snapshot, else with the oracle port.  Returns (seconds, kind)."""
    idx, workdir = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    if not _reference_importable():
        bank = tt.make_bank(eps_values=(0.01,))
        return oracle_forecast(bank)[0], "port"
    import contextlib
    import io

    sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))
    wd = os.path.join(workdir, f"w{idx}")
    os.makedirs(wd, exist_ok=True)
    os.chdir(wd)  # the reference creates ./cache and ./memory_cache at import
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        import cosmicfishpie.fishermatrix.cosmicfish as cff

        bank_dir = os.path.join(workdir, "bank")
        ext = tt.extfiles_for(bank_dir)
        t0 = time.perf_counter()
        out = []
        for obs, kw in ((["WL", "GCph"], dict(survey_name_photo=PHOTO_SPEC, survey_name_spectro=False, specs_dir=SPECDIR)),
                        (["GCsp"], dict(survey_name_spectro=SPECTRO_SPEC, survey_name_photo=False,
                                        spectro_Pk_k_samples=NK, spectro_Pk_mu_samples=NMU))):
            o = base_options(wd + "/", **kw)
            fm = cff.FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                                  specifications={}, observables=obs, extfiles=ext, cosmoModel="w0waCDM",
                                  surveyName="Euclid")
            out.append(fm.compute())
        dt = time.perf_counter() - t0
    return dt, "reference"


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the step on all host cores: one wave of
    `cores` independent single-threaded processes (how the reference parallelises -- Nautilus pool=N; a single
    Fisher is single-threaded), each computing ONE full forecast of the workload (no extrapolation: a forecast takes
    minutes, so the wave is the bounded sample and `--steps` only labels the line)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import multiprocessing as mp

    cores = max(1, min(os.cpu_count() or 1, 64))
    workdir = tempfile.mkdtemp(prefix="cfp_ref_")
    if _reference_importable():
        tt.write_bank(os.path.join(workdir, "bank"), eps_values=(0.01,))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_reference_worker, [(i, workdir) for i in range(cores)])
    wall = time.perf_counter() - t0
    kind = res[0][1]
    value = cores / wall
    sample = (f"one wave of {cores} concurrent single-threaded processes, each ONE full forecast (3x2pt 13+13 bins + GCsp "
              f"129x8193, 3PT) with the {'unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port'}: "
              f"wall {wall:.1f} s incl. interpreter start, slowest process {max(r[0] for r in res):.1f} s, "
              f"fastest {min(r[0] for r in res):.1f} s; --steps only labels the line")
    return {"impl": "reference", "metric": "fisher_matrices_per_s", "value": value, "unit": "Fisher/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / cores, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "Fisher/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Fisher/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


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
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix=f"cfp_bench_r{rank}_")
    bank = tt.make_bank(eps_values=(0.01,))
    ctx = _lib.default_context(local)

    def wmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- resident plans (host prep + upload outside the timed region of `value`) -------------------
    fm_ph = cdist.shard(make_photo_fm(bank, tmp), 0, 1)
    fm_sp = cdist.shard(make_spectro_fm(bank, tmp), 0, 1)
    fm_ph.compute()
    fm_sp.compute()
    pplan = fm_ph.photo_LSS._job.plan()
    splan = fm_sp._spectro_job.plan()
    ncell = NMU * NK
    nbins = pplan.Nl - 1
    cb, ce = cdist.split_range(ncell, rank, world)
    lb, le = cdist.split_range(nbins, rank, world)
    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream
    Fsp = cdist.device_view(splan.fisher_device_ptr(), (splan.Z, splan.npar, splan.npar), local)
    Fph = cdist.device_view(pplan.fisher_device_ptr(), (pplan.npar, pplan.npar), local)
    nph, nsp = Fph.numel(), Fsp.numel()
    fused = torch.zeros(nph + nsp, dtype=torch.float64, device="cuda")  # [F_photo | F_spectro]: ONE all-reduce per step
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    stream2 = torch.cuda.Stream()
    s2ptr = stream2.cuda_stream
    two_streams = os.environ.get("CFP_BENCH_ONE_STREAM", "0") != "1"

    def step_kernels(sharded=False):
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
                fused[:nph].copy_(Fph.view(-1))
                fused[nph:].copy_(Fsp.view(-1))

    def reduce_step(sharded):
        if sharded and world > 1:
            with torch.cuda.stream(stream):
                torch.distributed.all_reduce(fused)

    def step(sharded=False):
        step_kernels(sharded)
        reduce_step(sharded)

    use_graph = os.environ.get("CFP_BENCH_GRAPH", "1") != "0"
    graph_state = {}

    def captured(sharded):
        """The kernels of a step (both streams, and the packing of [F_photo | F_spectro]) as ONE CUDA graph: the nine
        launches of a forecast are issued from Python through ctypes, and once a forecast is split over 4-8 GPUs the
        kernels are shorter than the time it takes to launch them.  The NCCL all-reduce of the sharded step stays an
        eager call behind the replay.  Capture is thread-local (the NCCL watchdog thread queries events meanwhile).
        None when capture is not possible (then the loop stays eager)."""
        if not use_graph:
            return None
        if sharded not in graph_state:
            try:
                torch.cuda.synchronize()
                n0 = ctx.launches
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    step_kernels(sharded)
                graph_state[sharded] = (g, ctx.launches - n0)
            except Exception as exc:  # noqa: BLE001 -- report and fall back
                print(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {exc}); eager launches", file=sys.stderr)
                graph_state[sharded] = None
                torch.cuda.synchronize()
        return graph_state[sharded]

    def timed_loop(sharded, graph=False):
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
        cap = captured(sharded) if graph else None
        if graph and cap is None:
            return None, 0
        if cap is not None:
            with torch.cuda.stream(stream):
                for _ in range(max(3, args.warmup)):
                    cap[0].replay()
                    reduce_step(sharded)
            torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        n0 = ctx.launches
        with torch.cuda.stream(stream):
            for a, b in ev:
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
                a.record(stream)
                if cap is not None:
                    cap[0].replay()
                    reduce_step(sharded)
                else:
                    step(sharded)
                b.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        launched = cap[1] * args.steps if cap is not None else ctx.launches - n0
        return wmax(sum(a.elapsed_time(b) for a, b in ev)) / args.steps, launched

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # headline: a batch of independent forecasts, one per GPU per step (weak scaling, no collective)
    ms_eager, launches_timed = timed_loop(sharded=False)
    ms_graph, launches_graph = timed_loop(sharded=False, graph=True)
    graph_used = ms_graph is not None and ms_graph < ms_eager
    ms_per_step = ms_graph if graph_used else ms_eager
    if graph_used:
        launches_timed = launches_graph
    value = world * 1e3 / ms_per_step
    # what the timed steps computed is the forecast: compare the resident plans' results with the public-API run
    torch.cuda.synchronize()
    step_check = max(
        float(np.max(np.abs(Fph.cpu().numpy() - fm_ph.fisher_array)) / np.max(np.abs(fm_ph.fisher_array))),
        float(np.max(np.abs(Fsp.cpu().numpy().sum(axis=0) - fm_sp.fisher_array)) / np.max(np.abs(fm_sp.fisher_array))))
    if not step_check < 1e-12:
        raise RuntimeError(f"timed steps do not reproduce the API result (max rel diff {step_check:.3e})")
    # one forecast split over the GPUs (lattice cells / multipole bins) + ONE NCCL all-reduce of [F_photo | F_spectro]
    strong = None
    if world > 1:
        ms_strong_eager, _ = timed_loop(sharded=True)
        ms_strong_graph, _ = timed_loop(sharded=True, graph=True)
        ms_strong = min(ms_strong_eager, ms_strong_graph) if ms_strong_graph is not None else ms_strong_eager
        torch.cuda.synchronize()
        red = fused.cpu().numpy()
        chk_ph = float(np.max(np.abs(red[:nph].reshape(pplan.npar, pplan.npar) - fm_ph.fisher_array)) / np.max(np.abs(fm_ph.fisher_array)))
        chk_sp = float(np.max(np.abs(red[nph:].reshape(splan.Z, splan.npar, splan.npar).sum(axis=0) - fm_sp.fisher_array))
                       / np.max(np.abs(fm_sp.fisher_array)))
        if not max(chk_ph, chk_sp) < 1e-12:
            raise RuntimeError(f"sharded + all-reduced forecast differs from the single-GPU one ({chk_ph:.2e}, {chk_sp:.2e})")
        strong = {"value": 1e3 / ms_strong, "unit": "Fisher/s", "ms_per_step": ms_strong, "scaling": "strong",
                  "ms_per_step_eager_launches": ms_strong_eager, "ms_per_step_cuda_graph": ms_strong_graph,
                  "allreduced_vs_single_gpu_max_rel_diff": max(chk_ph, chk_sp),
                  "what": "ONE forecast per step split over the ranks (spectro lattice cells, photo multipole bins), "
                          "one fused NCCL all-reduce of [F_photo | F_spectro] inside the timed region; result asserted "
                          "equal to the single-GPU forecast"}
    splan.set_slice(cb, ce)
    pplan.set_slice(lb, le)

    # ---- dominant kernel alone: spectro_fisher_pairs_kernel ------------------------------------------
    kms, wms = [], []
    for _ in range(max(3, min(args.steps, 10))):
        with torch.cuda.stream(stream):
            flush.zero_()
        splan.run(materialize=0, stream=sptr)
        kms.append(splan.kernel_ms(2))
        wms.append(splan.kernel_ms(1))
    k_ms, w_ms = float(np.mean(kms)), float(np.mean(wms))
    pairs_ran = k_ms > 0
    if not pairs_ran:  # (a configuration the pair kernel does not take: the general kernel's time)
        k_ms, w_ms = float(splan.kernel_ms(0)), 0.0
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
    rows_slice = (ce + NK - 1) // NK - cb // NK
    nquad = (NK + 3) // 4
    warp_steps = splan.Z * nquad * (rows_slice + 2)  # (+1 pipeline-fill iteration per task, two tasks per column)
    flops_pairs = F_PAIR_STEP * warp_steps
    flops_weights = F_WEIGHT_CELL * splan.Z * rows_slice * NK
    peak_dfma = ctx.peak_fp64(0)
    peak_dmma = ctx.peak_fp64(1)
    achieved = flops_pairs / (k_ms * 1e-3) / 1e12
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s", "frac": achieved / peak_dfma,
            "traffic": TRAFFIC_BYTES if world == 1 else None, "kernel": "spectro_fisher_pairs_kernel", "kernel_ms": k_ms,
            "prepass_kernel": "spectro_weights_kernel", "prepass_ms": w_ms,
            "lattice_pass_frac": (flops_pairs + flops_weights) / ((k_ms + w_ms) * 1e-3) / 1e12 / peak_dfma,
            "peak_source": "measured live: cfp_peak_fp64 DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry)",
            "peak_dmma_tflops": peak_dmma, "flop_per_warp_step": F_PAIR_STEP, "warp_steps": warp_steps,
            "flop_per_cell_prepass": F_WEIGHT_CELL,
            "what": "FP64 operations the kernel executes (ncu-frozen count per warp step x warp steps) / CUDA-event "
                    "kernel time, against the DFMA rate measured live on this GPU.  The round-1 kernel executed 8.01 "
                    "GFLOP per launch for the same forecast; this one executes fewer (one division + a series instead "
                    "of two logarithms per stencil pair), so the fraction understates the speed-up"}
    # HBM view: the API-preserving mode writes derivs[npar][Z][cells] + veff[Z][cells] once
    with torch.cuda.stream(stream):
        flush.zero_()
    splan.run(materialize=_lib.MAT_DERIVS, stream=sptr)
    m_ms = splan.kernel_ms(0)
    bytes_alg = 8.0 * splan.Z * (ce - cb) * (splan.npar + 1)
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
        return ({"profile_run": True, "ms_per_step": ms_per_step, "kernel_ms": k_ms, "prepass_ms": w_ms,
                 "photo_kernel_ms": photo_ms, "roofline": roof, "ms_per_step_eager_launches": ms_eager,
                 "ms_per_step_cuda_graph": ms_graph, "strong_scaling": strong} if rank == 0 else None)

    # ---- the 5PT variant of the GCsp forecast (BASELINE config 4 says 5PT; the reference itself always runs 3PT) ----
    fm5 = cdist.shard(make_spectro_fm(bank, tmp, stencil="5PT"), 0, 1)
    fm5.compute()
    p5 = fm5._spectro_job.plan()
    t5 = []
    for _ in range(5):
        with torch.cuda.stream(stream):
            flush.zero_()
        p5.run(materialize=0, stream=sptr)
        t5.append(p5.kernel_ms(0))
    five_pt = {"gcsp_5pt_lattice_pass_ms": float(np.mean(t5)), "stencil_points": int(p5.S),
               "what": "GCsp 129x8193 with the 5-point stencil (45 stencil points instead of 23): device time of the "
                       "lattice pass (general kernel; oracle-pinned, the reference has no 5PT)"}
    fm5._spectro_job.close()

    # ---- end to end through the public API, from host tables -----------------------------------------
    from cosmicfishpie_b200 import cosmology as ccos

    def close_fm(f):
        job = f.photo_LSS._job if hasattr(f, "photo_LSS") else f._spectro_job
        job.close()  # (the fitted table bank stays resident on the device: cosmology._BANK_CACHE; clear_caches drops it)

    def bank_h2d(job):
        b = job.cosmo_set.bank()
        return getattr(b, "h2d_last", b.nbytes)  # 0 when the job found its bank resident

    def api_single(cold):
        if cold:
            ccos.clear_caches()  # tables -> FITPACK splines are re-fitted, as the reference does per Fisher
        a = cdist.shard(make_photo_fm(bank, tmp), 0, 1)  # every rank its own full forecast (weak scaling, like `value`)
        Fa = a.compute()
        b = cdist.shard(make_spectro_fm(bank, tmp), 0, 1)
        Fb = b.compute()
        tot = Fa + Fb
        h2d = (a.photo_LSS._job.plan().h2d_bytes + bank_h2d(a.photo_LSS._job)
               + b._spectro_job.plan().h2d_bytes + bank_h2d(b._spectro_job))
        d2h = 8 * (a.fisher_array.size + b.pk_fisher_per_bin.size)
        close_fm(a)
        close_fm(b)
        return tot, h2d, d2h

    def api_many(n):
        """n forecasts (2 n FisherMatrix objects) through the pipelined batch entry; per-forecast results summed."""
        fms = []
        for _ in range(n):
            fms.append(cdist.shard(make_photo_fm(bank, tmp), 0, 1))
            fms.append(cdist.shard(make_spectro_fm(bank, tmp), 0, 1))
        res = FisherMatrix.compute_many(fms)
        tot = [res[2 * i] + res[2 * i + 1] for i in range(n)]
        for f in fms:
            close_fm(f)
        return tot

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        out = fn(n)
        torch.cuda.synchronize()
        return wmax((time.perf_counter() - t0) / n), out

    n_e2e = max(2, min(args.steps, 5))
    api_single(True)
    t_cold, _ = timed(lambda n: [api_single(True) for _ in range(n)][-1], max(2, n_e2e // 2))
    api_single(False)
    t_single, (Ftot, h2d, d2h) = timed(lambda n: [api_single(False) for _ in range(n)][-1], n_e2e)
    n_many = max(8, args.steps)
    api_many(2)
    t_api, many = timed(api_many, n_many)
    many_check = float(np.max(np.abs(many[-1].fisher_matrix - Ftot.fisher_matrix)) / np.max(np.abs(Ftot.fisher_matrix)))
    if not many_check < 1e-13:
        raise RuntimeError(f"compute_many differs from compute ({many_check:.2e})")

    # ---- batched likelihoods (BASELINE.json: likelihood evals / s), each rank its own batch (weak) --------
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood

    def like_block(like, keys, allp, B, sigma, cosmo_cycle=False, shard=False):
        rng = np.random.default_rng(1234 + (0 if shard else rank))
        fid = np.array([allp[k] for k in keys])
        mat = fid * (1.0 + sigma * rng.standard_normal((B, len(keys))))
        if cosmo_cycle:  # every point its own cosmology of the bank: Omegam, h, sigma8 at fiducial x {0.99, 1, 1.01}, one at a time
            mat = np.hstack([np.tile([tt.FIDUCIAL[k] for k in ("Omegam", "h", "sigma8")], (B, 1)), mat])
            which = rng.integers(0, 3, size=B)
            mat[np.arange(B), which] *= rng.choice([0.99, 1.01], size=B)
            keys = ["Omegam", "h", "sigma8"] + list(keys)
        like.loglike_batch(mat, keys=keys, shard=shard)  # warm-up at the timed batch size (the memory pool grows once)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        ll = like.loglike_batch(mat, keys=keys, shard=shard)
        dt = wmax(time.perf_counter() - t0)
        total = B if shard else world * B
        return {"batch": B if shard else None, "batch_per_gpu": None if shard else B, "evals_per_s": total / dt,
                "device_ms_last_chunk": ctx.last_run_ms, "finite": bool(np.all(np.isfinite(ll)))}

    Lp = PhotometricLikelihood(cosmo_data=fm_ph)
    kp = [k for k in fm_ph.freeparams if k not in P7]
    Ls = SpectroLikelihood(cosmoFM_data=fm_sp)
    ks = [k for k in fm_sp.freeparams if k.startswith("lnbg")]
    Ll = SpectroLikelihood(cosmoFM_data=fm_sp, leg_flag="legendre")
    like = {
        "photo_3x2pt_13bins": like_block(Lp, kp, fm_ph.allparams, 16384, 0.02),
        "photo_3x2pt_13bins_cosmology_per_point": like_block(Lp, kp, fm_ph.allparams, 8192, 0.02, cosmo_cycle=True),
        "spectro_wedges_129x8193": like_block(Ls, ks, fm_sp.allparams, 16384, 0.01),
        "spectro_legendre_129x8193": like_block(Ll, ks, fm_sp.allparams, 16384, 0.01),
        "scaling": "weak",
        "what": "loglike_batch through the public API (host prep + H2D + kernels + D2H).  The spectroscopic nuisance-only "
                "batches collapse into groups that share a cosmology (moment kernels: one lattice pass per group and bin); "
                "`*_cosmology_per_point` / `spectro_*_single_points` are batches in which every point needs its own lattice "
                "pass or its own Limber integrals",
    }
    # every point its own spectroscopic lattice pass: 64 points that differ in sigma_p/sigma_v-free quantities cannot be
    # grouped, emulate with distinct (cosmology, bias) pairs drawn so that no two points share a block
    rng = np.random.default_rng(7 + rank)
    Bs = 256
    mat = np.tile([fm_sp.allparams[k] for k in ["h"] + ks], (Bs, 1)).astype(float)
    mat[:, 0] *= rng.choice([0.99, 1.0, 1.01], size=Bs)
    mat[:, 1:] += 0.01 * rng.standard_normal((Bs, len(ks)))
    os.environ["CFP_NO_MOMENTS"] = "1"  # the direct kernels: what a batch costs when nothing can be shared
    Ls.loglike_batch(mat, keys=["h"] + ks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Ls.loglike_batch(mat, keys=["h"] + ks)
    dt = wmax(time.perf_counter() - t0)
    os.environ.pop("CFP_NO_MOMENTS")
    like["spectro_wedges_129x8193_direct_kernels"] = {"batch_per_gpu": Bs, "evals_per_s": world * Bs / dt,
                                                      "what": "moment path switched off: every point and bin is one lattice pass"}
    # every point its OWN cosmology, continuous in all seven parameters (bank.TaylorEmulator): what a Nautilus chain
    # over cosmology asks for.  The bank of a batch is combined from the tabulated rows on the device; background
    # splines, no-wiggle spectra and velocity moments of the batch's cosmologies are fitted there in one call each.
    try:
        from cosmicfishpie_b200 import cosmology as ccos_
        from cosmicfishpie_b200.bank import TaylorEmulator

        emu_ext = dict(ext_dict(bank), emulator=TaylorEmulator(bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01))
        fe_ph = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                             options=base_options(tmp, survey_name_photo=PHOTO_SPEC, survey_name_spectro=False),
                             observables=["WL", "GCph"], extfiles=emu_ext)
        fe_ph.compute()
        fe_sp = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                             options=base_options(tmp, survey_name_spectro=SPECTRO_SPEC, survey_name_photo=False,
                                                  spectro_Pk_k_samples=NK, spectro_Pk_mu_samples=NMU),
                             observables=["GCsp"], extfiles=emu_ext)
        fe_sp.compute()
        rng = np.random.default_rng(99 + rank)
        for tag, L, fm_e, Be in (("photo_3x2pt_13bins", PhotometricLikelihood(cosmo_data=fe_ph), fe_ph, 2048),
                                 ("spectro_wedges_129x8193", SpectroLikelihood(cosmoFM_data=fe_sp), fe_sp, 512)):
            keys_e = list(P7) + [k for k in fm_e.freeparams if k not in P7][:4]
            fid_e = np.array([fm_e.allparams[k] for k in keys_e])
            mat_e = np.tile(fid_e, (Be, 1))
            u = 0.008 * rng.uniform(-1, 1, (Be, 7))
            mat_e[:, :7] = np.where(fid_e[:7] != 0, fid_e[:7] * (1 + u), u)
            mat_e[:, 7:] *= 1 + 0.01 * rng.standard_normal((Be, len(keys_e) - 7))
            L.loglike_batch(mat_e[:64], keys=keys_e)
            ccos_._BANK_CACHE.clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ll = L.loglike_batch(mat_e, keys=keys_e)
            dt = wmax(time.perf_counter() - t0)
            like[tag + "_emulated_cosmology_per_point"] = {
                "batch_per_gpu": Be, "evals_per_s": world * Be / dt, "finite": bool(np.all(np.isfinite(ll))),
                "what": "every point a distinct cosmology (all 7 parameters vary, second-order Taylor emulator over the "
                        "table bank); bank rows combined on the device (cfp_bank_combine_rows), 1-D fits / no-wiggle / "
                        "velocity moments of the batch fitted on the device"}
            fm_e.release_device()
    except Exception as exc:  # noqa: BLE001 -- an extra line, never the reason the bench fails
        like["emulated_cosmology_per_point_error"] = f"{type(exc).__name__}: {exc}"
    # one batch sharded over the ranks through the library (loglike_batch(shard=True)): slices + ONE all-gather
    like["sharded"] = {
        "photo_3x2pt_13bins": like_block(Lp, kp, fm_ph.allparams, 16384 * max(1, world), 0.02, shard=True),
        "spectro_wedges_129x8193": like_block(Ls, ks, fm_sp.allparams, 131072, 0.01, shard=True),
        "scaling": "strong", "what": "ONE batch split over the ranks, all-gather of the log-likelihoods inside the timed region"}

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
        if world == 1 and not args.no_cpu:
            t_cpu, detail = oracle_forecast(bank)
            cpu = {"value": 1.0 / t_cpu, "unit": "Fisher/s", "cores": 1, "kind": "port",
                   "sample": "ONE full forecast of the workload with the oracle port on one host core, measured (no "
                             "extrapolation): 3x2pt 13+13-bin Fisher %.1f s + GCsp Fisher on the 129x8193 lattice %.1f s"
                             % (detail["photo_s"], detail["spectro_s"]),
                   "likelihood": oracle_likelihoods(bank)}
            t5c, _ = oracle_forecast(bank, nmu=17, nk=1025, stencil="5PT", photo=False)
            five_pt["cpu_port_gcsp_5pt_17x1025_s"] = t5c
        line = {
            "metric": "fisher_matrices_per_s", "value": value, "unit": "Fisher/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2_flush_between_steps": True, "streams_per_step": 2 if two_streams else 1,
                       "cuda_graph": bool(graph_used),
                       "partition": "a batch of independent forecasts, one per GPU per step, no collective "
                                    "(strong_scaling: one forecast sharded over the GPUs + one fused NCCL all-reduce)"},
            "ms_per_step_eager_launches": ms_eager, "ms_per_step_cuda_graph": ms_graph,
            "strong_scaling_value": None if strong is None else strong["value"],
            "strong_scaling": strong,
            "clocks": clk,
            "e2e": {"value": world / t_api, "unit": "Fisher/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "public API: a batch of %d forecasts through FisherMatrix.compute_many (FisherMatrix(...) per "
                            "probe, per-job host prep, H2D of the job arrays, kernels, D2H, file export; host prep of forecast "
                            "i+1 overlaps the kernels of forecast i).  The fitted table bank of the forecast's cosmologies "
                            "stays resident on the device and the 1-D spline facades of the cached Boltzmann tables are "
                            "reused across steps (north_star: Boltzmann input is computed once and cached); e2e_cold "
                            "uploads the raw tables and re-fits everything on the device every step" % n_many,
                    "s_per_step": t_api},
            "e2e_single": {"value": world / t_single, "unit": "Fisher/s", "s_per_step": t_single,
                           "what": "one blocking compute() per probe (the reference's call pattern)"},
            "e2e_cold": {"value": world / t_cold, "unit": "Fisher/s", "s_per_step": t_cold,
                         "what": "as e2e_single but every cache cleared each step (raw tables -> H2D -> spline fits, distance integrals, "
                                 "no-wiggle spectra and velocity moments on the device -> bank), i.e. what "
                                 "the reference redoes for every Fisher matrix"},
            "cabi_e2e": {"value": world / t_cabi, "unit": "Fisher/s", "s_per_step": t_cabi,
                         "what": "cfp_*_plan_create(host buffers) + run + fetch, table bank resident"},
            "likelihood": like, "five_point_stencil": five_pt,
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the one-core CPU baseline (about two minutes)")
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

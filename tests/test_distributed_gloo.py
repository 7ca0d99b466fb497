"""CPU, world_size 2, gloo: the multi-GPU host logic (slice assignment, all-reduce of the partial Fisher
matrices) with the device plans replaced by deterministic stand-ins."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pattern(n):
    """A fixed well-conditioned symmetric matrix (the result container inverts the Fisher)."""
    return 1.0 + np.arange(n)[:, None] + np.arange(n)[None, :] + 10.0 * n * np.eye(n)


class FakeSpectroPlan:
    """F[z] = sum over the slice of a fixed per-cell pattern: slices must tile the lattice exactly."""

    def __init__(self, job):
        self.Z, self.npar = len(job.zmids), len(job.parnames)
        self.n = job.kgrid.size * job.mugrid.size
        self.begin, self.end = 0, self.n
        self.slices = []

    def set_slice(self, b, e):
        self.begin, self.end = b, e
        self.slices.append((b, e))

    def run(self, materialize=0, stream=0):
        pass

    def fetch(self, lnp=False, derivs=False):
        c = np.arange(self.begin, self.end, dtype=float)
        base = np.array([np.sum(np.cos(0.001 * c * (z + 1))) for z in range(self.Z)])
        F = base[:, None, None] * _pattern(self.npar)[None]
        return F, None, None, None


class FakePhotoPlan:
    def __init__(self, job):
        self.Nl, self.npar = len(job.grid.ell), job.stw.shape[0]
        self.begin, self.end = 0, self.Nl - 1

    def set_slice(self, b, e):
        self.begin, self.end = b, e

    def run(self, stream=0):
        pass

    def fetch(self, cls=False, sqrtp=False):
        l = np.arange(self.begin, self.end, dtype=float)
        return np.sum(np.sin(0.1 * l + 1.0)) * _pattern(self.npar), None, None


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import base_options
    from cosmicfishpie_b200 import photo, spectro, toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    bank = tt.make_bank(eps_values=(0.01,))
    ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS),
           "folder_paramnames": list(tt.COSMO_KEYS), "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}
    plans = {}
    spectro.SpectroJob.plan = lambda self: plans.setdefault(id(self), FakeSpectroPlan(self))
    photo.PhotoJob.plan = lambda self: plans.setdefault(id(self), FakePhotoPlan(self))
    o = base_options(os.path.join(tmp, f"r{rank}"), spectro_Pk_k_samples=65, spectro_Pk_mu_samples=9)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=o,
                      observables=["GCsp"], extfiles=ext)
    F = fm.compute()
    sp_plan = list(plans.values())[0]
    o2 = base_options(os.path.join(tmp, f"p{rank}"), survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                      survey_name_spectro=False, ell_sampling=21)
    fm2 = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01}, options=o2, observables=["GCph"],
                       extfiles=ext)
    F2 = fm2.compute()
    q.put((rank, F.fisher_matrix, sp_plan.slices[0], F2.fisher_matrix))
    dist.destroy_process_group()


def test_two_rank_sharded_compute_sums_to_single_rank(tmp_path):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, F0, s0, P0), (r1, F1, s1, P1) = res
    n = 65 * 9
    assert s0 == (0, (n + 1) // 2) and s1 == ((n + 1) // 2, n)  # contiguous, balanced, exact cover
    assert np.array_equal(F0, F1) and np.array_equal(P0, P1)  # every rank holds the reduced result
    # single-process expectation of the stand-in plans over the whole range
    c = np.arange(n, dtype=float)
    npar = F0.shape[0]
    full = sum(np.sum(np.cos(0.001 * c * (z + 1))) for z in range(4)) * _pattern(npar)
    assert np.allclose(F0, full, rtol=1e-12)
    nb = 21 - 1
    assert np.allclose(P0, np.sum(np.sin(0.1 * np.arange(nb) + 1.0)) * _pattern(P0.shape[0]), rtol=1e-12)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosmicfishpie_b200.likelihood import Likelihood

    class Toy(Likelihood):
        """chi2 = sum of squares of the row: every rank must only ever see its own slice."""

        def __init__(self):
            self.cosmo_theory = None
            self.seen = []

        def compute_data(self):
            return None

        def compute_theory(self, p):
            return None

        def compute_chi2(self, t):
            return 0.0

        def chi2_batch(self, dicts):
            raise AssertionError("the array path is the one under test")

        def chi2_matrix(self, names, matrix):
            self.seen.append(matrix.shape[0])
            return np.sum(np.asarray(matrix) ** 2, axis=1)

    like = Toy()
    mat = np.arange(7 * 3, dtype=float).reshape(7, 3)  # 7 points over 2 ranks: slices of 4 and 3
    ll = like.loglike_batch(mat, keys=["a", "b", "c"], shard=True)
    q.put((rank, ll, like.seen))
    dist.destroy_process_group()


def test_two_rank_sharded_loglike_batch_gathers_every_point():
    """loglike_batch(shard=True): contiguous slices per rank, one all-gather, every rank gets the full vector."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = -0.5 * np.sum(np.arange(21, dtype=float).reshape(7, 3) ** 2, axis=1)
    for rank, ll, seen in res:
        assert np.array_equal(ll, expect)
        assert seen == ([4] if rank == 0 else [3])

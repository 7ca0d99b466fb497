"""GPU tests of the runtime pieces around the kernels: the stream-ordered device arena, page-locked table rows,
device-resident likelihood data, the compact (z-map) galaxy-bias upload and lattice slices.  Each one checks that the
faster path returns exactly what the plain path returns."""

import numpy as np
import pytest

from conftest import base_options

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]


@pytest.fixture(scope="module")
def spectro_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                      options=base_options(tmp_path_factory.mktemp("rt")), observables=["GCsp"], extfiles=extfiles)
    fm.compute()
    return fm


def test_bank_from_pinned_rows_equals_bank_from_blob(spectro_fm):
    from cosmicfishpie_b200 import _lib

    cs = spectro_fm._spectro_job.cosmo_set
    ctx = cs.ctx
    rows = [c.bank_row(cs.slots) for c in cs.cosmos[:3]]
    a = _lib.Bank(ctx, rows)
    b = _lib.Bank.from_rows(ctx, [_lib.BankRow(ctx, r) for r in rows])
    rng = np.random.default_rng(3)
    z, k = rng.uniform(0.0, 2.5, 500), np.exp(rng.uniform(np.log(2e-4), np.log(20.0), 500))
    for c in range(3):
        for slot in cs.slots:
            assert np.array_equal(a.eval(c, slot, z, k), b.eval(c, slot, z, k))
    assert a.nbytes == b.nbytes
    a.close()
    b.close()


def test_pinned_blocks_are_recycled(spectro_fm):
    from cosmicfishpie_b200 import _lib

    ctx = spectro_fm._spectro_job.cosmo_set.ctx
    p = _lib.PinnedBuffer(ctx, 12345)
    addr = p.ptr.value
    p.array[:] = 1.5
    p.close()
    q = _lib.PinnedBuffer(ctx, 12345)
    assert q.ptr.value == addr and q.array.shape == (12345,)
    q.close()


def test_arena_and_plain_allocator_agree(spectro_fm, monkeypatch):
    """CFP_NO_POOL=1 makes a context use cudaMalloc/cudaFree: same numbers as the pooled default."""
    from cosmicfishpie_b200 import _lib

    cs = spectro_fm._spectro_job.cosmo_set
    rows = [cs.cosmos[0].bank_row(cs.slots)]
    z, k = np.linspace(0.1, 2.0, 64), np.geomspace(1e-3, 5.0, 64)
    want = _lib.Bank(cs.ctx, rows).eval(0, cs.slots[0], z, k)
    monkeypatch.setenv("CFP_NO_POOL", "1")
    plain = _lib.Context(cs.ctx.device)
    monkeypatch.delenv("CFP_NO_POOL")
    bank = _lib.Bank(plain, rows)
    assert np.array_equal(bank.eval(0, cs.slots[0], z, k), want)
    bank.close()
    plain.close()


def test_plan_create_destroy_cycles_do_not_grow_memory(spectro_fm):
    """The arena hands the same blocks back: after a warm-up cycle, further create/run/destroy cycles leave the
    device's free memory unchanged."""
    import torch

    job = spectro_fm._spectro_job

    def cycle():
        job.close()
        job.plan().run()
        F, _, _, _ = job.plan().fetch()
        return F

    F0 = cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(5):
        F = cycle()
    torch.cuda.synchronize()
    assert np.array_equal(F, F0)
    assert abs(torch.cuda.mem_get_info()[0] - free0) <= 4 << 20


def test_wedge_chi2_with_device_resident_data(spectro_fm):
    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200.likelihood import SpectroLikelihood

    like = SpectroLikelihood(cosmoFM_data=spectro_fm)
    job = spectro_fm._spectro_job
    plan = job.plan()
    shot = np.zeros((plan.S, plan.Z))
    host = plan.chi2(shot, data=like.data_wedges)
    dev = plan.chi2(shot, data_dev=_lib.DeviceArray(plan.ctx, like.data_wedges))
    assert np.array_equal(host, dev)
    assert host[0] < 1e-15 and np.all(host[1:] > host[0])  # the fiducial point reproduces its own data


def test_photo_plan_zmap_equals_dense_bias(extfiles, tmp_path):
    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from cosmicfishpie_b200.photo import PhotoJob

    o = base_options(tmp_path, survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                      observables=["WL", "GCph"], extfiles=extfiles)
    names = ["b1", "b4", "b10", "AIA"]
    rng = np.random.default_rng(5)
    mat = np.array([fm.allparams[n] for n in names]) * (1 + 0.03 * rng.standard_normal((6, len(names))))
    job = PhotoJob.from_matrix(fm, names, mat)
    assert job.zmap is not None and job.bias.shape == (6, len(fm.specs["binrange_GCph"]))  # one bias vector per point
    compact = job.run().fetch(cls=True)[1]
    g = job.grid
    dense_bias = job.dense_bias()  # [B, Nb, Nz]
    cs = job.cosmo_set
    dense = _lib.PhotoPlan(cs.ctx, cs.bank(), cosmo_of_s=job.cosmo_of_s, ell=g.ell, z=g.z, dz=g.dz, chi=job.chi,
                           hub=job.hub, wlpref=job.wlpref, kind=g.kind, ngal=job.ngal, bias=dense_bias, ia=job.ia_window(),
                           lmin=g.lmin, lmax=g.lmax, noise=g.noise, sqrtfsky=g.sqrtfsky, stw=job.stw, stden=job.stden,
                           kmax=g.kmax, tab_p=g.tab_p)
    dense.run()
    # (the compact job evaluates the IA window on the device: CUDA's pow against numpy's, an ulp apart)
    assert np.allclose(dense.fetch(cls=True)[1], compact, rtol=1e-13, atol=0.0)
    dense.close()
    job.close()


def test_spectro_slices_sum_to_full_fisher(spectro_fm):
    """Any partition of the lattice into cell ranges (what the ranks of a multi-GPU run get) sums to the full
    matrix to rounding; ragged cuts inside a mu row included."""
    job = spectro_fm._spectro_job
    plan = job.plan()
    ncell = plan.Nmu * plan.Nk
    plan.set_slice(0, ncell)
    plan.run()
    full = plan.fetch()[0]
    cuts = [0, plan.Nk // 3, 5 * plan.Nk + 7, ncell - 11, ncell]
    acc = np.zeros_like(full)
    for a, b in zip(cuts[:-1], cuts[1:]):
        plan.set_slice(a, b)
        plan.run()
        acc += plan.fetch()[0]
    plan.set_slice(0, ncell)
    assert np.max(np.abs(acc - full)) <= 1e-11 * np.max(np.abs(full))


def test_two_plans_on_two_streams(spectro_fm, extfiles, tmp_path):
    """Plans of one context running concurrently on the context's two streams: each fetch waits for its own plan and
    returns what the serial run returns."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    o = base_options(tmp_path, survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False,
                     ell_sampling=30)
    photo_fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7[:3]}, options=o,
                            observables=["WL", "GCph"], extfiles=extfiles)
    photo_fm.compute()
    pp, sp = photo_fm.photo_LSS._job.plan(), spectro_fm._spectro_job.plan()
    assert pp.ctx is sp.ctx
    pp.run()
    sp.run()
    want_p, want_s = pp.fetch()[0], sp.fetch()[0]
    for _ in range(5):
        sp.run()                          # the long lattice kernel first, on the context stream
        pp.run(stream=pp.ctx.aux_stream)  # the short photometric kernels on the second stream
        got_p = pp.fetch()[0]             # must wait for pp, not for whichever plan ran last on the context
        got_s = sp.fetch()[0]
        assert np.array_equal(got_p, want_p) and np.array_equal(got_s, want_s)

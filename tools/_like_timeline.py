import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cosmicfishpie_b200 import toy_tables as tt, photo as photo_mod
from cosmicfishpie_b200.likelihood import PhotometricLikelihood
tmp = tempfile.mkdtemp()
fm = bench.make_photo_fm(tt.make_bank(eps_values=(0.01,)), tmp); fm.compute()
like = PhotometricLikelihood(cosmo_data=fm)
keys = [k for k in fm.freeparams if k not in bench.P7]
fid = np.array([fm.allparams[k] for k in keys])
mat = fid * (1.0 + 0.02 * np.random.default_rng(1).standard_normal((16384, len(keys))))
like.loglike_batch(mat, keys=keys)
off, n, w = like._blocks()
for rep in range(2):
    T0 = time.perf_counter(); rows = []; pending = None
    for a in range(0, 16384, 4096):
        t0 = time.perf_counter()
        job = photo_mod.PhotoJob.from_matrix(like.cosmo_theory, keys, mat[a:a + 4096]); t1 = time.perf_counter()
        plan = job.plan(); t2 = time.perf_counter()
        plan.chi2_launch(off, n, w, data=like._data_full, stream=plan.ctx.aux_stream); t3 = time.perf_counter()
        if pending is not None:
            pending.plan().chi2_collect(); pending.close()
        t4 = time.perf_counter()
        pending = job
        rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    pending.plan().chi2_collect(); pending.close()
    print("total %.1f ms; per chunk from_matrix / plan / launch / collect(prev): " % (1e3 * (time.perf_counter() - T0)) + " | ".join(" ".join("%.1f" % (1e3 * x) for x in r) for r in rows))
# discriminate: jobs prebuilt; between launch and collect either sleep or busy numpy work
jobs = [photo_mod.PhotoJob.from_matrix(like.cosmo_theory, keys, mat[a:a + 4096]) for a in range(0, 16384, 4096)]
for mode in ("sleep", "numpy", "none"):
    jobs = [photo_mod.PhotoJob.from_matrix(like.cosmo_theory, keys, mat[a:a + 4096]) for a in range(0, 16384, 4096)]
    rows = []
    for job in jobs:
        plan = job.plan(); t2 = time.perf_counter()
        plan.chi2_launch(off, n, w, data=like._data_full, stream=plan.ctx.aux_stream); t3 = time.perf_counter()
        if mode == "sleep":
            time.sleep(0.009)
        elif mode == "numpy":
            x = np.random.rand(4096, 200)
            te = time.perf_counter()
            while time.perf_counter() - te < 0.009:
                x = np.sqrt(x * 1.0001)
        t4 = time.perf_counter()
        plan.chi2_collect(); t5 = time.perf_counter()
        job.close()
        rows.append((t3 - t2, t4 - t3, t5 - t4))
    print(mode, "launch / between / collect:", " | ".join(" ".join("%.1f" % (1e3 * x) for x in r) for r in rows))

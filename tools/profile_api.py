"""cProfile of the public-API forecast (`FisherMatrix(...).compute()`, warm process caches) on the GPU box:
where the host side of `bench.py`'s e2e number goes.  Usage: python tools/profile_api.py [n_iter] [out.txt]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    out = sys.argv[2] if len(sys.argv) > 2 else None
    tmp = tempfile.mkdtemp(prefix="cfp_prof_")
    bank = tt.make_bank(eps_values=(0.01,))

    def step(which):
        res = []
        if which in ("photo", "both"):
            a = bench.make_photo_fm(bank, tmp)
            res.append(a.compute())
            a.photo_LSS._job.close()
            a.photo_LSS._job.cosmo_set.release_bank()
        if which in ("spectro", "both"):
            b = bench.make_spectro_fm(bank, tmp)
            res.append(b.compute())
            b._spectro_job.close()
            b._spectro_job.cosmo_set.release_bank()
        return res

    step("both")
    step("both")
    buf = io.StringIO()
    # wall-clock phases (no profiler): constructor | host preparation of the job | H2D + kernels + D2H | file export
    for which, make in (("photo", bench.make_photo_fm), ("spectro", bench.make_spectro_fm)):
        acc = {"ctor": 0.0, "host_prep": 0.0, "device": 0.0, "export": 0.0, "close": 0.0}
        for _ in range(n):
            t0 = time.perf_counter()
            fm = make(bank, tmp)
            t1 = time.perf_counter()
            fm.compute()
            t2 = time.perf_counter()
            job = fm.photo_LSS._job if which == "photo" else fm._spectro_job
            job.close()
            job.cosmo_set.release_bank()
            t3 = time.perf_counter()
            acc["ctor"] += t1 - t0
            acc["host_prep"] += fm.timings["host_prep_s"]
            acc["device"] += fm.timings["device_s"]
            acc["export"] += (t2 - t1) - fm.timings["compute_s"]
            acc["close"] += t3 - t2
        buf.write(f"==== {which} phases, ms per forecast: " + ", ".join(f"{k} {1e3 * v / n:.2f}" for k, v in acc.items()) + "\n")
    for which in ("photo", "spectro"):
        t0 = time.perf_counter()
        for _ in range(n):
            step(which)
        dt = (time.perf_counter() - t0) / n
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(n):
            step(which)
        pr.disable()
        buf.write(f"==== {which}: {dt * 1e3:.2f} ms per forecast (unprofiled), n={n}\n")
        ps = pstats.Stats(pr, stream=buf).strip_dirs().sort_stats("cumulative")
        ps.print_stats(45)
        ps.sort_stats("tottime").print_stats(30)
    txt = buf.getvalue()
    if out:
        open(out, "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()

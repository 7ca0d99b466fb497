"""Minimal workload for ncu / timing of the device-side preparation (csrc/cfp_prep.cu): COLD forecasts of the bench
workload -- every cache cleared, so each forecast fits its table bank, background splines, no-wiggle spectra and
velocity moments on the device again.  Usage: python tools/prep_driver.py [n_forecasts] [--host]"""
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
if "--host" in sys.argv:
    os.environ["CFP_HOST_PREP"] = "1"

import bench  # noqa: E402
from cosmicfishpie_b200 import cosmology as cz  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
    bank = tt.make_bank()
    tmp = tempfile.mkdtemp()
    times = []
    for i in range(n + 1):
        cz.clear_caches()
        t0 = time.perf_counter()
        a = bench.make_photo_fm(bank, tmp)
        a.compute()
        t1 = time.perf_counter()
        b = bench.make_spectro_fm(bank, tmp)
        b.compute()
        t2 = time.perf_counter()
        if i:
            times.append((t1 - t0, t2 - t1))
        a.release_device()
        b.release_device()
    ph = sum(t[0] for t in times) / len(times) * 1e3
    sp = sum(t[1] for t in times) / len(times) * 1e3
    mode = "host (scipy) preparation" if os.environ.get("CFP_HOST_PREP") else "device preparation"
    print(f"cold forecast, {mode}: photo {ph:.2f} ms + spectro {sp:.2f} ms = {ph + sp:.2f} ms")
    if "--profile" in sys.argv:
        import cProfile
        import pstats

        def cold():
            cz.clear_caches()
            a = bench.make_photo_fm(bank, tmp)
            a.compute()
            b = bench.make_spectro_fm(bank, tmp)
            b.compute()
            a.release_device()
            b.release_device()

        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            cold()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(40)


if __name__ == "__main__":
    main()

"""A/B of FisherMatrix.compute_many on the bench workload: result files written by the helper thread (default) or
synchronously (CFP_SYNC_EXPORT=1).  Usage: python tools/ab_many.py [n_forecasts] [repeats]"""
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402
from cosmicfishpie_b200.fishermatrix import FisherMatrix  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    bank = tt.make_bank()
    tmp = tempfile.mkdtemp()

    def many():
        fms = []
        for _ in range(n):
            fms.append(bench.make_photo_fm(bank, tmp))
            fms.append(bench.make_spectro_fm(bank, tmp))
        t0 = time.perf_counter()
        FisherMatrix.compute_many(fms)
        return (time.perf_counter() - t0) / n * 1e3

    many()
    for mode in ("0", "1", "0", "1"):
        os.environ["CFP_SYNC_EXPORT"] = mode
        ts = sorted(many() for _ in range(reps))
        print(f"CFP_SYNC_EXPORT={mode}: compute_many only (constructors outside) {ts[0]:.3f} .. {ts[len(ts) // 2]:.3f} .. {ts[-1]:.3f} ms per forecast")


if __name__ == "__main__":
    main()

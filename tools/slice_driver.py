"""Minimal workload for an ncu launch list of ONE rank's share of a sharded forecast: the resident plans of the bench
workload with the slices rank 0 of N would get (no collective).  Usage: python tools/slice_driver.py N [steps]"""
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import distributed as cdist  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bank = tt.make_bank()
    tmp = tempfile.mkdtemp()
    a = cdist.shard(bench.make_photo_fm(bank, tmp), 0, 1)
    b = cdist.shard(bench.make_spectro_fm(bank, tmp), 0, 1)
    a.compute()
    b.compute()
    pplan, splan = a.photo_LSS._job.plan(), b._spectro_job.plan()
    cb, ce = cdist.split_range(bench.NMU * bench.NK, 0, world)
    lb, le = cdist.split_range(pplan.Nl - 1, 0, world)
    splan.set_slice(cb, ce)
    pplan.set_slice(lb, le)
    for _ in range(steps):
        pplan.run()
        splan.run(materialize=0)
    splan.fetch()
    print("pairs ms", splan.kernel_ms(2), "weights ms", splan.kernel_ms(1), "photo", [pplan.kernel_ms(i) for i in range(3)])


if __name__ == "__main__":
    main()

"""A/B timing of the photometric chain of the bench forecast (3x2pt 13+13 bins, 47 stencil points x 100 multipoles):
device time of the Limber table, C_ell and Fisher kernels (CUDA events inside the library, L2 flushed between runs),
for the libcfp_b200.so / environment currently in effect (CFP_CLS_WARPS, CFP_CLS_LCH).
Usage: python tools/ab_photo.py [label] [n_runs]"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "lib"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    fm = bench.make_photo_fm(tt.make_bank(eps_values=(0.01,)), tempfile.mkdtemp(prefix="cfp_abp_"))
    fm.compute()
    plan = fm.photo_LSS._job.plan()
    stream = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(n + 5):
        with torch.cuda.stream(stream):
            flush.zero_()
        plan.run(stream=stream.cuda_stream)
        torch.cuda.synchronize()
        if i >= 5:
            ms.append([plan.kernel_ms(k) for k in range(3)])
    ms = np.mean(ms, axis=0)
    F = plan.fetch()[0]
    print(f"{label}: limber {ms[0]:.4f} cls {ms[1]:.4f} fisher {ms[2]:.4f} ms   trace {np.trace(F):.10e}")


if __name__ == "__main__":
    main()

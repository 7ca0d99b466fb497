#!/usr/bin/env python
"""Print parity (vs reference goldens) and timing of every golden case on the GPU box."""
import json, os, sys, tempfile, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import base_options, golden
from cosmicfishpie_b200 import toy_tables as tt
from cosmicfishpie_b200.fishermatrix import FisherMatrix
import test_spectro_gpu as ts, test_photo_gpu as tp

bank = tt.make_bank()
ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
       "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}
tmp = tempfile.mkdtemp()
rows = []
for mod, cases in ((ts, ts.CASES), (tp, tp.CASES)):
    for name in cases:
        g = golden(name)
        for rep in range(2):
            t0 = time.time(); fm, F = mod.run_case(name, ext, tmp); dt = time.time() - t0
        A, B = F.fisher_matrix, g["fisher"]; nz = B != 0
        rows.append(dict(case=name, npar=int(A.shape[0]), fisher_max_rel=float(np.max(np.abs(A[nz]-B[nz])/np.abs(B[nz]))),
                         sigma_max_rel=float(np.max(np.abs(F.get_confidence_bounds()-g["sigma"])/g["sigma"])),
                         wall_s=dt, ref_wall_s=float(g["wall_s"]), **{k: float(v) for k, v in fm.timings.items()}))
        print(json.dumps(rows[-1]))

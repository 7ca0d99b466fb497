import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The CPU suite exercises the host logic (job assembly against fake plans, world-size-2 gloo): without a device the
    # batched preparation of the cosmology input (csrc/cfp_prep.cu) cannot run, so these tests take the host fits.
    # On the GPU box the default (device preparation) is what every test goes through.
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        os.environ.setdefault("CFP_HOST_PREP", "1")


@pytest.fixture(scope="session")
def toy_bank():
    """In-memory toy table bank (fiducial +- 1% per cosmological parameter), identical to the
    text files the goldens were generated from (tools/make_goldens.py)."""
    from cosmicfishpie_b200 import toy_tables as tt

    return tt.make_bank(eps_values=(0.01,))


@pytest.fixture(scope="session")
def extfiles(toy_bank):
    from cosmicfishpie_b200 import toy_tables as tt

    return {"tables": toy_bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS),
            "folder_paramnames": list(tt.COSMO_KEYS), "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / scale) if scale > 0 else float(np.max(np.abs(a)))


def elem_rel_err(a, b):
    """max over elements of |a-b|/|b| (elements where b == 0 must be exactly 0 in a)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    nz = b != 0
    assert np.all(a[~nz] == 0), "non-zero where the reference is exactly zero"
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))) if nz.any() else 0.0


def base_options(tmp, **kw):
    o = {"accuracy": 1, "feedback": 0, "code": "external", "outroot": "t", "results_dir": str(tmp),
         "survey_name": "Euclid", "cosmo_model": "w0waCDM", "derivatives": "3PT",
         "survey_name_spectro": "Euclid-Spectroscopic-ISTF-Pessimistic", "survey_name_photo": False}
    o.update(kw)
    return o

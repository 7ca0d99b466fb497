"""CPU: the C-ABI shared library loads and exports every symbol include/cfp_b200.h declares
(no compute calls: there is no GPU in the build container)."""
import os
import re

import pytest

from conftest import REPO


def declared_symbols():
    src = open(os.path.join(REPO, "include", "cfp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cosmicfishpie_b200 import _lib

    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cfp_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert lib.cfp_abi_version() == _lib.ABI_VERSION


def test_header_constants_match_binding():
    from cosmicfishpie_b200 import _lib

    src = open(os.path.join(REPO, "include", "cfp_b200.h")).read()
    enum = re.search(r"enum \{\s*CFP_SP_Z = 0,(.*?)CFP_SP_NSCAL", src, flags=re.S).group(1)
    enum = re.sub(r"/\*.*?\*/", "", enum, flags=re.S)
    n = len([t for t in enum.split(",") if t.strip()]) + 1
    assert n == _lib.SP_NSCAL
    assert int(re.search(r"#define CFP_BANK_DESC (\d+)", src).group(1)) == _lib.BANK_DESC
    assert int(re.search(r"#define CFP_NTAB (\d+)", src).group(1)) == _lib.NTAB


def test_no_gpu_means_loud_failure():
    """Without a usable CUDA device the product raises; it never falls back to a CPU path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cosmicfishpie_b200 import _lib

    with pytest.raises(_lib.CfpError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "cosmicfishpie_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f

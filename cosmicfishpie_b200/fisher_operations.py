"""Fisher-matrix post-processing on the device, batched.

The reference post-processes one matrix at a time on the CPU: `fisher_matrix.__add__` (analysis/fisher_matrix.py:462-548),
`get_confidence_bounds` (:705-733), `fisher_operations.marginalise` / `marginalise_over`
(analysis/fisher_operations.py:177-272).  A GPU produces forecasts by the thousand (batches of fiducials / survey
specifications, `FisherMatrix.compute_many`); these functions take the whole batch -- arrays [B][n][n] or lists of
`fisher_matrix` objects -- and run ONE launch per operation (cfp_fisher_batch_post / cfp_fisher_batch_add).  The
single-matrix functions of the reference's module are the B = 1 case and return `fisher_matrix` objects."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .fisher_matrix import confidence_coefficient, fisher_matrix


def _ctx(device=None):
    return _lib.default_context(device)


def _stack(fishers) -> np.ndarray:
    if isinstance(fishers, np.ndarray):
        return np.atleast_3d(fishers) if fishers.ndim == 3 else fishers[None]
    return np.array([f.fisher_matrix for f in fishers])


def batch_confidence_bounds(fishers, confidence_level: float = 0.6827, device=None) -> np.ndarray:
    """Marginal 1-sigma bounds [B][n] of a batch of Fisher matrices: coef * sqrt(diag(F^-1))."""
    sigma, _, _ = _ctx(device).fisher_batch_post(_stack(fishers), None, confidence_coefficient(confidence_level))
    return sigma


def batch_marginalise(fishers, param_names: Sequence[str], keep: Sequence[str], confidence_level: float = 0.6827, device=None):
    """For every matrix of the batch (all with parameters `param_names`): marginal bounds [B][n], the Fisher matrix
    marginalised onto `keep` (in that order) [B][m][m], and sqrt(det) of it [B] -- with keep = ("w0", "wa") the
    dark-energy figure of merit."""
    names = list(param_names)
    for k in keep:
        if k not in names:
            raise ValueError("Error, parameter " + str(k) + " is not in a parameter of fisher_matrix")
    idx = [names.index(k) for k in keep]
    return _ctx(device).fisher_batch_post(_stack(fishers), idx, confidence_coefficient(confidence_level))


def batch_figure_of_merit(fishers, param_names: Sequence[str], pair=("w0", "wa"), device=None) -> np.ndarray:
    """sqrt(det F_marginal(pair)) of every matrix: 1 / sqrt(det Cov[pair, pair])."""
    return batch_marginalise(fishers, param_names, list(pair), device=device)[2]


def batch_add(first: List[fisher_matrix], second: List[fisher_matrix], device=None) -> List[fisher_matrix]:
    """Element-wise sums first[b] + second[b] over the union of their parameters (every first[b] shares one parameter
    list, every second[b] another): one launch for the whole batch."""
    if len(first) != len(second) or not first:
        raise ValueError("batch_add needs two non-empty lists of equal length")
    na, nb = list(first[0].param_names), list(second[0].param_names)
    for f in first:
        if list(f.param_names) != na:
            raise ValueError("the matrices of a batch must share their parameter list")
    for f in second:
        if list(f.param_names) != nb:
            raise ValueError("the matrices of a batch must share their parameter list")
    names = na + [p for p in nb if p not in na]
    out = _ctx(device).fisher_batch_add(_stack(first), _stack(second), [names.index(p) for p in na],
                                        [names.index(p) for p in nb], len(names))
    res = []
    for b, (fa, fb) in enumerate(zip(first, second)):
        latex, fid = list(names), np.zeros(len(names))
        for f in (fa, fb):
            for j, p in enumerate(f.param_names):
                latex[names.index(p)] = f.param_names_latex[j]
                fid[names.index(p)] = f.param_fiducial[j]
        res.append(fisher_matrix(fisher_matrix=out[b], param_names=names, param_names_latex=latex, fiducial=fid))
    return res


# ---- the reference's single-matrix interface (analysis/fisher_operations.py:177-272): the B = 1 case -------------
def marginalise(fisher: fisher_matrix, names: Iterable[str], update_names: bool = True, device=None) -> fisher_matrix:
    """Marginalise over all parameters but `names`; the result has the parameters of `names`, in that order."""
    if not isinstance(fisher, fisher_matrix):
        raise ValueError("Error, input fisher_matrix is not a fisher_matrix")
    names = list(names)
    _, marg, _ = batch_marginalise([fisher], fisher.param_names, names, device=device)
    idx = [fisher.param_names_dict[n] - 1 for n in names]
    out = fisher_matrix(fisher_matrix=marg[0], param_names=names,
                        param_names_latex=[fisher.param_names_latex[i] for i in idx],
                        fiducial=[fisher.param_fiducial[i] for i in idx], name=fisher.name)
    if update_names:
        out.name = fisher.name + "_marginal"
    out.path, out.indir = getattr(fisher, "path", None), getattr(fisher, "indir", None)
    return out


def marginalise_over(fisher: fisher_matrix, names: Iterable[str], device=None) -> fisher_matrix:
    """Marginalise over the parameters in `names`; the result keeps all the others."""
    if not isinstance(fisher, fisher_matrix):
        raise ValueError("Error, input fisher_matrix is not a fisher_matrix")
    names = list(names)
    for n in names:
        if n not in fisher.param_names_dict:
            raise ValueError("Error, parameter " + str(n) + " is not in a parameter of fisher_matrix")
    return marginalise(fisher, [p for p in fisher.param_names if p not in names], device=device)

"""Result container returned by `FisherMatrix.compute()`.

A compact stand-in for the part of the reference's `cosmicfishpie.analysis.fisher_matrix.fisher_matrix`
(analysis/fisher_matrix.py:36-861) that the hot path and its tests touch: construction from a matrix or
from the exported `<root>.txt` + `<root>.paramnames`, deletion of all-zero rows/columns (:261-268), `tril`
symmetrisation (:616-629), inverse, marginal 1-sigma bounds (:705-733) and `__add__` (:462-548).  The
plotting / derived-parameter machinery of the reference's analysis package is out of scope.
"""
from __future__ import annotations

import copy
import os

import numpy as np
from scipy.special import erfinv


def confidence_coefficient(confidence_level: float) -> float:
    return np.sqrt(2.0) * erfinv(confidence_level)


class fisher_matrix:
    def __init__(self, fisher_matrix=None, param_names=None, param_names_latex=None, fiducial=None,
                 file_name=None, name=""):
        if fisher_matrix is None and file_name is None:
            raise ValueError("Error in initializing the Fisher matrix: fisher_matrix and file_name are both None.")
        self.file_name = file_name
        self.name = name
        names_from_file = None
        if fisher_matrix is None:
            self.fisher_matrix = np.loadtxt(file_name)
            self.path = os.path.abspath(file_name)
            self.name = os.path.splitext(os.path.basename(file_name))[0]
            self.indir = os.path.dirname(self.path)
            pn = os.path.splitext(self.path)[0] + ".paramnames"
            if os.path.isfile(pn):
                names_from_file = self._read_paramnames(pn)
        else:
            self.fisher_matrix = np.array(fisher_matrix, dtype=float)
        if self.fisher_matrix.ndim == 0:
            self.fisher_matrix = self.fisher_matrix.reshape(1, 1)
        elif self.fisher_matrix.ndim == 1:
            self.fisher_matrix = self.fisher_matrix.reshape(1, -1)
        m = self.fisher_matrix
        if not (np.abs(m - m.T) <= 1e-06 + 1e-03 * np.abs(m.T)).all():  # == not np.allclose(m, m.T, rtol=1e-3, atol=1e-6)
            m = np.tril(m) + np.triu(m, 1)
        zr = np.where(~m.any(axis=1))[0]
        zc = np.where(~m.any(axis=0))[0]
        if zc.any() and zc.any():  # (sic) reference condition, analysis/fisher_matrix.py:263
            m = np.delete(np.delete(m, zr, axis=0), zc, axis=1)
        self.fisher_matrix = np.copy(m)
        self.num_params = m.shape[0]
        if param_names is None and names_from_file is not None and len(names_from_file[0]) == self.num_params:
            self.param_names, self.param_names_latex, fid = names_from_file
            self.param_fiducial = np.array(fid)
        elif param_names is None:
            self.param_names = ["p" + str(i + 1) for i in range(self.num_params)]
            self.param_names_latex = list(self.param_names)
            self.param_fiducial = np.zeros(self.num_params)
        else:
            self.param_names = list(param_names)  # (names are strings: a new list is a deep copy)
            self.param_names_latex = list(param_names)
            self.param_fiducial = np.zeros(self.num_params)
        if param_names_latex is not None:
            self.param_names_latex = list(param_names_latex)
        if fiducial is not None:
            self.param_fiducial = np.array(fiducial, dtype=float)
        if len(self.param_names) != self.num_params:
            raise ValueError("The input param_names has not " + str(self.num_params) + " elements.")
        self.param_names_dict = {}
        for i, n in enumerate(self.param_names):
            self.param_names_dict[i + 1] = n
            self.param_names_dict[n] = i + 1

    # eigen-decomposition and inverse (analysis/fisher_matrix.py computes them in the constructor): on first access
    def _eigen(self):
        if "_eig" not in self.__dict__:
            self._eig = np.linalg.eigh(self.fisher_matrix)
        return self._eig

    @property
    def fisher_eigenvalues(self):
        return self._eigen()[0]

    @property
    def fisher_eigenvectors(self):
        return self._eigen()[1]

    @property
    def fisher_matrix_inv(self):
        if "_inv" not in self.__dict__:
            self._inv = self.inverse_fisher_matrix()
        return self._inv

    @staticmethod
    def _read_paramnames(path):
        """`name    latex    fiducial` per line, four-space separated (latex may contain single spaces)."""
        names, latex, fid = [], [], []
        with open(path) as fh:
            for line in fh:
                if not line.strip() or line.lstrip().startswith("#"):
                    continue
                parts = [t.strip() for t in line.split("    ") if t.strip() != ""]
                names.append(parts[0])
                if len(parts) == 1:
                    latex.append(parts[0])
                    fid.append(0.0)
                elif len(parts) == 2:
                    try:
                        fid.append(float(parts[1]))
                        latex.append(parts[0])
                    except ValueError:
                        fid.append(0.0)
                        latex.append(parts[1])
                else:
                    latex.append(parts[1])
                    fid.append(float(parts[2]))
        return names, latex, fid

    # getters ---------------------------------------------------------------------------------
    def get_fisher_matrix(self):
        return self.fisher_matrix

    def get_fisher_inverse(self):
        return self.fisher_matrix_inv

    def get_param_names(self):
        return self.param_names

    def get_param_names_latex(self):
        return self.param_names_latex

    def get_param_fiducial(self):
        return self.param_fiducial

    def get_param_index(self, name):
        return self.param_names_dict[name] - 1

    def inverse_fisher_matrix(self):
        return np.linalg.inv(np.array(self.fisher_matrix))

    def determinant(self):
        return np.linalg.det(self.fisher_matrix)

    def get_confidence_bounds(self, confidence_level=0.6827, marginal=True, cache=False):
        if confidence_level < 0.0 or confidence_level > 1.0:
            raise ValueError("Invalid confidence level. Legal input is between 0 and 1.")
        inv = self.fisher_matrix_inv if cache else self.inverse_fisher_matrix()
        errors = np.sqrt(np.diagonal(inv)) if marginal else np.sqrt(1.0 / np.diagonal(self.fisher_matrix))
        return confidence_coefficient(confidence_level) * errors

    def __add__(self, other):
        """Sum over the union of parameters (reference: analysis/fisher_matrix.py:462-548)."""
        names = list(self.param_names) + [n for n in other.param_names if n not in self.param_names]
        n = len(names)
        out = np.zeros((n, n))
        fid = np.zeros(n)
        latex = list(names)
        for fm in (self, other):
            idx = [names.index(p) for p in fm.param_names]
            out[np.ix_(idx, idx)] += fm.fisher_matrix
            for j, p in zip(idx, range(fm.num_params)):
                fid[j] = fm.param_fiducial[p]
                latex[j] = fm.param_names_latex[p]
        return fisher_matrix(fisher_matrix=out, param_names=names, param_names_latex=latex, fiducial=fid)

    def save_to_file(self, file_name, simple_header=False, file_format=".txt"):
        root = os.path.splitext(file_name)[0]
        with open(root + file_format, "w") as fh:
            fh.write("# " + " ".join(self.param_names) + "\n")
            np.savetxt(fh, self.fisher_matrix, fmt="%.17e")
        with open(root + ".paramnames", "w") as fh:
            for n, l, f in zip(self.param_names, self.param_names_latex, self.param_fiducial):
                fh.write(f"{n}    {l}    {f:.6f}\n")

"""Table bank: one binary file instead of a directory tree of text files, and a cosmology-continuous emulator over
it.

The reference's `external` input is a directory per cosmology with eight text files each (cosmology.py:1249-1373), and
a parameter point that is not one of the tabulated +-eps cosmologies silently SNAPS to the nearest one
(cosmology.py:1375-1421) -- a likelihood can only move among the tabulated folders (photo_like.py:171-177 would call
the Boltzmann solver per point in `camb`/`class` mode, which is outside this package).

* `save_bank` / `load_bank`: the whole bank {folder: {table: array}} as ONE file: an 8-byte magic, a JSON header
  (folders, table names, shapes, byte offsets, free-form metadata) and the float64 payload, 64-byte aligned; loaded
  zero-copy through `numpy.memmap`.  `extfiles["tables"] = load_bank(path)` is all a FisherMatrix needs.
* `TaylorEmulator`: tables of an ARBITRARY cosmology from the fiducial and the +-eps folders of the bank, second order
  in every parameter (the diagonal of the Hessian; cross terms are not tabulated):
      T(theta) = T_0 + sum_p [ d_p (T_p+ - T_p-) / 2 + d_p^2 (T_p+ + T_p- - 2 T_0) / 2 ],   d_p = (theta_p - theta_p0) / h_p
  i.e. fixed WEIGHTS over the bank's folders (they sum to one; at a tabulated cosmology they are one-hot, so bank nodes
  are reproduced exactly).  With `extfiles["emulator"] = emulator` the cosmology facade builds an off-node cosmology from
  these tables instead of snapping, so likelihood batches move continuously in all cosmological parameters.
  Because interpolating-spline fitting is linear in the table values, the same weights applied to the bicubic
  coefficient rows already resident on the device give the emulated cosmology's bank row without any host fit:
  `cfp_bank_combine` (`_lib.Bank.combined`)."""
from __future__ import annotations

import json
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

MAGIC = b"CFPBANK1"
_ALIGN = 64


def save_bank(path: str, tables: Dict[str, Dict[str, np.ndarray]], meta: Optional[dict] = None) -> str:
    """Write {folder: {name: array}} as one binary bank file; returns `path`."""
    entries, off = [], 0
    for folder, t in tables.items():
        for name, a in t.items():
            a = np.ascontiguousarray(a, dtype=np.float64)
            entries.append({"folder": folder, "name": name, "shape": list(a.shape), "offset": off})
            off += (a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    header = json.dumps({"version": 1, "dtype": "<f8", "meta": meta or {}, "entries": entries}).encode()
    pad = (-(len(MAGIC) + 8 + len(header))) % _ALIGN
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(np.uint64(len(header) + pad).tobytes())
        fh.write(header + b" " * pad)
        base = fh.tell()
        for e in entries:
            a = np.ascontiguousarray(tables[e["folder"]][e["name"]], dtype="<f8")
            fh.seek(base + e["offset"])
            fh.write(a.tobytes())
        fh.truncate(base + off)
    return path


def load_bank(path: str, mmap: bool = True):
    """(tables, meta) of a bank file; arrays are read-only views of a memory map unless `mmap` is False."""
    with open(path, "rb") as fh:
        if fh.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path} is not a cosmicfishpie_b200 table bank")
        hlen = int(np.frombuffer(fh.read(8), dtype=np.uint64)[0])
        header = json.loads(fh.read(hlen).decode())
        base = fh.tell()
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    tables: Dict[str, Dict[str, np.ndarray]] = {}
    for e in header["entries"]:
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        a = np.frombuffer(buf, dtype="<f8", count=n, offset=base + e["offset"]).reshape(e["shape"])
        tables.setdefault(e["folder"], {})[e["name"]] = a
    return tables, header.get("meta", {})


class LazyTables:
    """Dictionary-like view of an emulated cosmology's tables: `t[name]` forms the weighted sum on first access."""

    def __init__(self, emulator: "TaylorEmulator", weights: Dict[str, float]):
        self.emulator, self.weights, self._done = emulator, weights, {}
        self._wvec = emulator.weight_vector(weights)

    def __contains__(self, name):
        return name in self.emulator.tables[self.emulator.fiducial_folder]

    def keys(self):
        return self.emulator.tables[self.emulator.fiducial_folder].keys()

    def __getitem__(self, name):
        if name not in self._done:
            t0 = self.emulator.tables[self.emulator.fiducial_folder]
            if name in TaylorEmulator.GRID_KEYS:
                self._done[name] = np.asarray(t0[name])
            else:
                if name not in t0:
                    raise KeyError(name)
                first = np.asarray(t0[name])
                self._done[name] = self.emulator.combine(name, self._wvec if first.ndim == 1 else self.weights)
        return self._done[name]

    def column0(self, name):
        """First wavenumber column of a (z, k) table without forming the table."""
        if name in self._done:
            return np.asarray(self._done[name])[:, 0]
        return self.emulator.combine(name, self._wvec, column=0)


class TaylorEmulator:
    """Second-order (diagonal) Taylor emulator over a fiducial +- eps table bank; see the module docstring."""

    GRID_KEYS = ("z", "k")  # abscissae: identical in every folder, never combined

    def __init__(self, tables, fiducial: Dict[str, float], paramnames: Sequence[str], eps: float,
                 folder_paramnames: Optional[Sequence[str]] = None, fiducial_folder: str = "fiducial_eps_0"):
        from .cosmology import folder_for

        self.tables, self.fiducial, self.eps = tables, dict(fiducial), float(eps)
        self.paramnames = list(paramnames)
        self.fiducial_folder = fiducial_folder
        ext = {"paramnames": self.paramnames, "folder_paramnames": list(folder_paramnames or paramnames),
               "eps_values": [self.eps], "fiducial_folder": fiducial_folder}
        self.step, self.folders = {}, {}
        for p in self.paramnames:
            h = self.fiducial[p] * self.eps if self.fiducial[p] != 0.0 else self.eps  # derivatives.py:140-143
            self.step[p] = h
            pl = folder_for(dict(self.fiducial, **{p: self.fiducial[p] + h}), self.fiducial, ext)
            mn = folder_for(dict(self.fiducial, **{p: self.fiducial[p] - h}), self.fiducial, ext)
            if pl not in tables or mn not in tables:
                raise KeyError(f"the bank lacks the +-eps folders of {p}: {pl}, {mn}")
            self.folders[p] = (pl, mn)
        t0 = tables[fiducial_folder]
        for f, t in tables.items():
            for g in self.GRID_KEYS:
                if not np.array_equal(np.asarray(t[g]), np.asarray(t0[g])):
                    raise ValueError(f"folder {f} is tabulated on another {g} grid than the fiducial")

    def offsets(self, theta: Dict[str, float]) -> Dict[str, float]:
        """d_p = (theta_p - fiducial_p) / step_p."""
        out = {}
        for p in self.paramnames:
            d = (float(theta.get(p, self.fiducial[p])) - self.fiducial[p]) / self.step[p]
            r = round(d)
            out[p] = float(r) if abs(d - r) < 1e-9 and abs(r) <= 1 else d  # (a tabulated step, up to rounding)
        return out

    def weights(self, theta: Dict[str, float]) -> Dict[str, float]:
        """{folder: weight}; the weights sum to one."""
        w = {self.fiducial_folder: 1.0}
        for p, d in self.offsets(theta).items():
            if d == 0.0:
                continue
            pl, mn = self.folders[p]
            w[self.fiducial_folder] -= d * d
            w[pl] = w.get(pl, 0.0) + 0.5 * d + 0.5 * d * d
            w[mn] = w.get(mn, 0.0) - 0.5 * d + 0.5 * d * d
        return {f: x for f, x in w.items() if x != 0.0}

    def is_node(self, theta: Dict[str, float]) -> bool:
        """A tabulated cosmology: at most one parameter off the fiducial, by exactly one step."""
        d = [x for x in self.offsets(theta).values() if x != 0.0]
        return len(d) == 0 or (len(d) == 1 and abs(abs(d[0]) - 1.0) < 1e-9)

    def combine(self, name: str, weights: Dict[str, float], column: Optional[int] = None) -> np.ndarray:
        """sum_f weights[f] * tables[f][name] (one table; `column`: only that column of a (z, k) table).  Functions of z
        and single columns go through a stacked copy [folder][z] made once: a batch of thousands of emulated
        cosmologies forms three of them per cosmology."""
        first = np.asarray(self.tables[self.fiducial_folder][name])
        if first.ndim == 1 or column is not None:
            key = (name, column)
            stacks = self.__dict__.setdefault("_stacks", {})
            st = stacks.get(key)
            if st is None:
                st = stacks[key] = np.stack([np.asarray(self.tables[f][name], dtype=float) if column is None
                                             else np.asarray(self.tables[f][name], dtype=float)[:, column]
                                             for f in self.folder_order])
            return self.weight_vector(weights) @ st
        acc = None
        for folder, x in weights.items():
            term = x * np.asarray(self.tables[folder][name], dtype=float)
            acc = term if acc is None else acc + term
        return acc

    def weight_vector(self, weights) -> np.ndarray:
        """The weights {folder: x} as a dense vector over `folder_order` (a vector passes through)."""
        if isinstance(weights, np.ndarray):
            return weights
        col = self.__dict__.get("_col_of_folder")
        if col is None:
            col = self.__dict__["_col_of_folder"] = {f: i for i, f in enumerate(self.folder_order)}
        w = np.zeros(len(col))
        for folder, x in weights.items():
            w[col[folder]] = x
        return w

    def tables_at(self, theta: Dict[str, float], weights: Optional[Dict[str, float]] = None) -> "LazyTables":
        """The tables of cosmology `theta`: the weighted sum of the bank's folders, each table formed on first access
        (a facade whose (z, k) surfaces are combined on the device never forms them on the host)."""
        return LazyTables(self, self.weights(theta) if weights is None else weights)

    @property
    def folder_order(self):
        """The bank's folders in a fixed order: the columns of `weight_matrix` / rows of `base_bank`."""
        return list(self.tables)

    def base_bank(self, ctx, slots, table_of_slot):
        """Device bank of the tabulated cosmologies (rows in `folder_order`), fitted on the device once per (context,
        table selection) and kept: the rows every emulated cosmology is combined from (`cfp_bank_combine_rows`).
        `table_of_slot`: {bank slot: table name}.  None when the folders are not all on one grid."""
        from . import _lib

        key = (id(ctx), tuple(sorted(slots)))
        cache = self.__dict__.setdefault("_base_banks", {})
        hit = cache.get(key)
        if hit is not None and hit[0] is ctx and hit[1].handle:
            return hit[1]
        t0 = self.tables[self.fiducial_folder]
        z, k = np.asarray(t0["z"], float).ravel(), np.asarray(t0["k"], float).ravel()
        rows = []
        for folder in self.folder_order:
            row = [None] * _lib.NTAB
            for slot in slots:
                tab = self.tables[folder].get(table_of_slot[slot])
                if tab is None or np.shape(tab) != (z.size, k.size):
                    return None
                row[slot] = tab
            rows.append(row)
        bank = _lib.Bank.fit(ctx, z, np.tile(k, (len(rows), 1)), rows, [[1.0] * _lib.NTAB] * len(rows))
        cache[key] = (ctx, bank)
        return bank

    def weight_matrix(self, thetas: Iterable[Dict[str, float]], folder_order: Sequence[str]) -> np.ndarray:
        """[M][len(folder_order)] weights for `cfp_bank_combine`, columns in the order of the bank's rows."""
        thetas = list(thetas)
        W = np.zeros((len(thetas), len(folder_order)))
        col = {f: i for i, f in enumerate(folder_order)}
        for m, th in enumerate(thetas):
            for f, x in self.weights(th).items():
                W[m, col[f]] = x
        return W

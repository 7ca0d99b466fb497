"""Stand-in for the `nautilus-sampler` package (not installable offline): the part of its interface that the
reference's likelihood package imports and that this package's sampler driver calls.  Tools and tests only.

`Prior` records parameters and their distributions; `Sampler` is a deterministic importance sampler from the prior box
(uniform / frozen scipy distributions): enough to drive `loglike` / `loglike_batch` end to end, write a chain and an
evidence, and to check that the vectorised likelihood entry is what gets called.  It is NOT a nested sampler."""
import numpy as np


class Prior:
    def __init__(self):
        self.keys = []
        self.dists = []

    def add_parameter(self, key, dist=None):
        self.keys.append(key)
        self.dists.append(dist)

    @property
    def dimensionality(self):
        return len(self.keys)

    def sample(self, n, rng):
        cols = []
        for d in self.dists:
            if isinstance(d, tuple):
                cols.append(rng.uniform(d[0], d[1], size=n))
            elif hasattr(d, "rvs"):
                cols.append(np.asarray(d.rvs(size=n, random_state=rng)))
            else:
                cols.append(rng.uniform(0.0, 1.0, size=n))
        return np.stack(cols, axis=1)


class Sampler:
    def __init__(self, prior, likelihood, n_live=200, n_networks=1, n_batch=100, pool=None, pass_dict=True,
                 filepath=None, resume=True, likelihood_kwargs=None, vectorized=False, seed=0, **_):
        self.prior, self.likelihood = prior, likelihood
        self.n_live, self.n_batch, self.pool = int(n_live), int(n_batch), pool
        self.pass_dict, self.vectorized = pass_dict, vectorized
        self.kw = dict(likelihood_kwargs or {})
        self.filepath = filepath
        self.rng = np.random.default_rng(seed)
        self.points = np.empty((0, prior.dimensionality))
        self.log_l = np.empty(0)
        self.n_calls = 0

    def _evaluate(self, pts):
        self.n_calls += 1
        if self.vectorized:
            arg = {k: pts[:, i] for i, k in enumerate(self.prior.keys)} if self.pass_dict else pts
            return np.asarray(self.likelihood(arg, **self.kw), dtype=float)
        out = []
        for p in pts:
            arg = dict(zip(self.prior.keys, p)) if self.pass_dict else p
            out.append(float(self.likelihood(arg, **self.kw)))
        return np.array(out)

    def run(self, verbose=False, discard_exploration=False, n_batches=3, **_):
        for _ in range(n_batches):
            pts = self.prior.sample(max(self.n_live, self.n_batch), self.rng)
            self.points = np.vstack([self.points, pts])
            self.log_l = np.concatenate([self.log_l, self._evaluate(pts)])
        if self.filepath:
            np.save(self.filepath + ".npy", np.column_stack([self.points, self.log_l]))
        return True

    @property
    def log_z(self):
        m = np.max(self.log_l)
        return float(m + np.log(np.mean(np.exp(self.log_l - m))))

    def evidence(self):
        return self.log_z

    def posterior(self):
        m = np.max(self.log_l)
        w = np.exp(self.log_l - m)
        return self.points, np.log(w / np.sum(w)), self.log_l

"""Import stub so the reference's likelihood package imports without nautilus-sampler (tools only)."""


class Prior:
    def __init__(self):
        self.keys = []

    def add_parameter(self, key, dist=None):
        self.keys.append(key)


class Sampler:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("nautilus is not installed; this is an import stub")

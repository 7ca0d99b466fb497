"""CPU oracle: a numpy/scipy restatement of cosmicfishpie's observable -> Fisher / likelihood path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  The product
(``cosmicfishpie_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked in ``tests/test_oracle_golden.py``
against golden vectors produced by running the UNMODIFIED reference in the build container
(``tools/make_goldens.py`` -> ``tests/golden/*.npz``; the reference's own test goldens need
the ``symbolic_pofk`` backend that is not installable here).

Third-party arithmetic on the path (not under /root/reference) is reused as the same
third-party calls, exactly as the reference makes them: scipy FITPACK
(``RectBivariateSpline``, ``InterpolatedUnivariateSpline``, ``interp1d``),
``scipy.signal.savgol_filter``, ``scipy.integrate.simpson/trapezoid``, ``scipy.special.erf``,
``numpy.linalg.pinv/det/inv``.

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""

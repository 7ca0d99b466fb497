"""Host-side pieces of the device preparation (cosmicfishpie_b200/cosmology.py): the Savitzky-Golay filter as the linear
operator cfp_bank_nowiggle applies, FITPACK's interpolation knots."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline


def test_savgol_operator_is_scipys_filter():
    from scipy.signal import savgol_filter

    from cosmicfishpie_b200.cosmology import savgol_operator

    rng = np.random.default_rng(3)
    x = np.cumsum(rng.normal(size=800)) * 0.05 + np.linspace(5, -3, 800)
    for win, order in ((101, 3), (100, 3), (57, 3), (58, 2), (31, 2), (5, 3)):
        taps, off, EL, ER = savgol_operator(win, order)
        ne, wc = EL.shape
        y = np.empty_like(x)
        for s in range(x.size):
            if s < ne:
                y[s] = EL[s] @ x[:wc]
            elif s >= x.size - ne:
                y[s] = ER[s - (x.size - ne)] @ x[x.size - wc:]
            else:
                y[s] = taps @ x[s + off:s + off + win]
        assert np.max(np.abs(y - savgol_filter(x, win, order))) < 1e-13, (win, order)


def test_nak_knots_are_fitpacks():
    from cosmicfishpie_b200.cosmology import nak_knots

    rng = np.random.default_rng(0)
    for n in (5, 6, 17, 100):
        x = np.sort(rng.uniform(0, 3, n))
        t = InterpolatedUnivariateSpline(x, np.sin(x))._eval_args[0]
        assert np.array_equal(nak_knots(x), t)

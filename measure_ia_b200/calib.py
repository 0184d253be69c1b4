"""Exact bin thresholds, calibrated against numpy on this host.

The reference bins a pair with (``src/measureia/measure_w_box_jk.py:420-434``, ``measure_m_box_jk.py:443-460``)

    ind_r  = floor(log10(r) / dlog - log10(r_bins[0]) / dlog)           dlog = (log10(r_max) - log10(r_min)) / n_r
    ind_pi = floor(Pi / dpi - pi_bins[0] / dpi)                          dpi  = (pi_bins[-1] - pi_bins[0]) / n_pi
    ind_mu = floor(mu / dmu - mu_r_bins[0] / dmu)                        dmu  = 2.0 / n_pi

where ``log10`` is whatever kernel numpy dispatches to on this CPU (SVML on AVX-512 hosts: it differs from libm and
from CUDA's log10 in the last ulp).  Each index is a monotone step function of its argument, so it is fully described
by the smallest double at which each step happens.  We find those doubles by bisection *evaluating the reference's own
numpy expression*, and the GPU then only compares against them: the GPU's bin assignment is bit-identical to what the
reference computes on this machine, without evaluating a single logarithm or division on the device.

The separation enters the reference as ``r = sqrt(s)`` with ``s`` the sum of squares; ``sqrt`` is correctly rounded
and monotone, so thresholds on ``r`` are converted to thresholds on ``s`` (again by search with ``np.sqrt``) and the
device never takes the square root for binning either.
"""
import numpy as np

_I64 = np.int64


def _to_ord(x):
	"""Order-preserving map float64 -> int64."""
	b = np.asarray(x, dtype=np.float64).view(_I64)
	return np.where(b >= 0, b, np.int64(-2 ** 63) - b)


def _from_ord(k):
	k = np.asarray(k, dtype=_I64)
	b = np.where(k >= 0, k, np.int64(-2 ** 63) - k)
	return b.view(np.float64)


def first_true(pred, lo, hi):
	"""Smallest double x in (lo, hi] with pred(x) True, for a vectorised monotone predicate.

	lo / hi: arrays with pred(lo) False and pred(hi) True element-wise."""
	lo = np.array(lo, dtype=np.float64)
	hi = np.array(hi, dtype=np.float64)
	if not (np.all(~pred(lo)) and np.all(pred(hi))):
		raise ValueError("threshold bracket does not straddle the step")
	a, b = _to_ord(lo).copy(), _to_ord(hi).copy()
	while np.any(b - a > 1):
		mid = a + (b - a) // 2
		t = pred(_from_ord(mid))
		b = np.where(t, mid, b)
		a = np.where(t, a, mid)
	return _from_ord(b)


def _is_clean_step(pred, thr, halfwidth=64):
	"""pred is False on the `halfwidth` doubles below thr and True on thr and the doubles above."""
	k = _to_ord(thr)
	offs = np.arange(-halfwidth, halfwidth + 1, dtype=_I64)
	ok = True
	for kk in np.atleast_1d(k):
		vals = pred(_from_ord(kk + offs))
		ok &= bool(np.all(~vals[:halfwidth]) and np.all(vals[halfwidth:]))
	return ok


def _sqrt_threshold(t):
	"""Smallest double s with np.sqrt(s) >= t (t > 0)."""
	t = np.asarray(t, dtype=np.float64)
	guess = t * t
	lo = guess * (1 - 1e-12) - 5e-324
	hi = guess * (1 + 1e-12) + 5e-324
	return first_true(lambda s: np.sqrt(s) >= t, lo, hi)


def r_thresholds(r_min, r_max, n_r, r_bins):
	"""(thr_r[0..n_r], thr_s[0..n_r], clean): thresholds on r and on s = r^2 (see module docstring).

	thr[0] / thr[n_r] are the range mask ``r >= r_bins[0]`` / ``r < r_bins[-1]``; thr[b] for 0 < b < n_r is the
	smallest r whose floor-log index is >= b."""
	dlog = (np.log10(r_max) - np.log10(r_min)) / n_r
	c0 = np.log10(r_bins[0]) / dlog
	thr = np.empty(n_r + 1)
	thr[0], thr[n_r] = r_bins[0], r_bins[-1]
	clean = True
	if n_r > 1:
		b = np.arange(1, n_r, dtype=np.float64)
		edges = np.asarray(r_bins[1:-1], dtype=np.float64)

		def pred(x):
			with np.errstate(all="ignore"):
				return np.floor(np.log10(x) / dlog - c0) >= (b if x.shape == b.shape else b[:, None])

		thr[1:n_r] = first_true(pred, edges * (1 - 1e-9), edges * (1 + 1e-9))
		for i in range(1, n_r):
			bb = float(i)
			clean &= _is_clean_step(lambda x: np.floor(np.log10(x) / dlog - c0) >= bb, thr[i])
	if not np.all(np.diff(thr) > 0):
		raise ValueError("radial bin thresholds are not increasing; check separation_limits / num_bins_r")
	return thr, _sqrt_threshold(thr), clean


def linear_thresholds(lo, width, n, edges, first, last):
	"""Thresholds of floor(x / width - lo / width) on x (used for Pi and mu_r).

	edges: nominal bin edges (n + 1); first / last: values stored at thr[0] / thr[n] (range mask, or -inf / +inf)."""
	c0 = lo / width
	thr = np.empty(n + 1)
	thr[0], thr[n] = first, last
	if n > 1:
		b = np.arange(1, n, dtype=np.float64)
		e = np.asarray(edges[1:-1], dtype=np.float64)
		pad = 1e-9 * max(abs(edges[0]), abs(edges[-1]), 1e-300)

		def pred(x):
			return np.floor(x / width - c0) >= b

		thr[1:n] = first_true(pred, e - pad, e + pad)
	if not np.all(np.diff(thr[np.isfinite(thr)]) > 0):
		raise ValueError("second-axis bin thresholds are not increasing")
	return thr


def pi_thresholds(pi_bins, n_pi):
	width = (pi_bins[-1] - pi_bins[0]) / n_pi  # measure_w_box_jk.py:368
	return linear_thresholds(pi_bins[0], width, n_pi, pi_bins, pi_bins[0], pi_bins[-1])


def mu_thresholds(mu_r_bins, n_pi):
	width = 2.0 / n_pi  # measure_m_box_jk.py:385
	return linear_thresholds(mu_r_bins[0], width, n_pi, mu_r_bins, -np.inf, np.inf)


def rp_cut_threshold(rp_cut):
	"""Largest s with sqrt(s) <= rp_cut, so that ``r_p > rp_cut``  <=>  ``s > threshold`` (measure_m_box_jk.py:444)."""
	if rp_cut is None or rp_cut <= 0.0:
		return 0.0 if (rp_cut is None or rp_cut == 0.0) else -1.0
	t = np.float64(rp_cut)
	first_above = first_true(lambda s: np.sqrt(s) > t, np.array([t * t * (1 - 1e-12)]), np.array([t * t * (1 + 1e-12)]))
	return float(np.nextafter(first_above[0], -np.inf))

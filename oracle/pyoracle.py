"""CPU oracle: restatement of the reference's periodic-box measurement.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product (``measure_ia_b200/``) never does.  Parity status: PINNED -- validated against the
unmodified reference run under ``oracle/ref_shims`` (tests/test_oracle_vs_reference.py) and against the reference's own
golden HDF5 outputs (tests/test_golden_hdf5.py, fixtures decoded into tests/golden/ by oracle/make_golden.py).

The pair loop itself is C (oracle.c, loaded through ctypes); everything around it is numpy written in the
reference's evaluation order.  Citations are ``file:line`` under ``/root/reference/src/measureia``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Params(ctypes.Structure):
	_fields_ = [
		("geom", ctypes.c_int), ("n_r", ctypes.c_int), ("n_2", ctypes.c_int),
		("r_min", ctypes.c_double), ("r_max", ctypes.c_double),
		("r_edge_lo", ctypes.c_double), ("r_edge_hi", ctypes.c_double),
		("lo2", ctypes.c_double), ("hi2", ctypes.c_double),
		("boxsize", ctypes.c_double), ("half_box", ctypes.c_double),
		("periodic", ctypes.c_int), ("los", ctypes.c_int),
		("rp_cut", ctypes.c_double), ("num_box", ctypes.c_int), ("two_R", ctypes.c_double),
	]


def build(force=False):
	"""Compile oracle.c -> oracle/liboracle.so (building the checker is not using it)."""
	so = os.path.join(_HERE, "liboracle.so")
	src = os.path.join(_HERE, "oracle.c")
	if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
		subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
	return so


def _lib():
	global _LIB
	if _LIB is None:
		_LIB = ctypes.CDLL(build())
		_LIB.oracle_paircount.restype = ctypes.c_int
		_LIB.oracle_max_threads.restype = ctypes.c_int
	return _LIB


def max_threads():
	"""Host threads this process may use: the scheduler affinity mask, NOT OpenMP's default (torchrun exports
	OMP_NUM_THREADS=1 to its workers, which would make the CPU arm single-threaded).  oracle.c passes the count in an
	explicit ``num_threads`` clause, so the environment variable does not cap it."""
	try:
		n = len(os.sched_getaffinity(0))
	except AttributeError:
		n = os.cpu_count() or 1
	return max(1, int(n))


# --------------------------------------------------------------------------------------------------------------------
# host-side restatements
# --------------------------------------------------------------------------------------------------------------------
def make_bins(separation_limits, num_bins_r, num_bins_pi, pi_max, boxsize):
	"""measure_IA_base.py:155-167."""
	r_min, r_max = separation_limits
	r_bins = np.logspace(np.log10(r_min), np.log10(r_max), num_bins_r + 1)
	if pi_max is None:
		pi_max = boxsize / 2.0
	pi_bins = np.linspace(-pi_max, pi_max, num_bins_pi + 1)
	mu_r_bins = np.linspace(-1, 1, num_bins_pi + 1)
	return r_bins, pi_bins, mu_r_bins


def prepare(data, masks=None, ellipticity="distortion"):
	"""Input preparation repeated at the top of every reference variant, e.g. measure_w_box_jk.py:322-364."""
	if masks is None:
		pos, pos_s = data["Position"], data["Position_shape_sample"]
		axis_v, q = data["Axis_Direction"], data["q"]
		w, w_s = data["weight"], data["weight_shape_sample"]
	else:
		masks = dict(masks)
		n_p, n_s = len(data["Position"]), len(data["Position_shape_sample"])
		pos = data["Position"][masks["Position"]]
		pos_s = data["Position_shape_sample"][masks["Position_shape_sample"]]
		axis_v = data["Axis_Direction"][masks["Axis_Direction"]]
		q = data["q"][masks["q"]]
		if "weight" not in masks:  # :338-342 -- selects the FIRST sum(mask) weights (reference quirk)
			m = np.ones(n_p, dtype=bool)
			m[int(np.sum(masks["Position"])):n_p] = 0
			masks["weight"] = m
		if "weight_shape_sample" not in masks:  # :343-347
			m = np.ones(n_s, dtype=bool)
			m[int(np.sum(masks["Position_shape_sample"])):n_s] = 0
			masks["weight_shape_sample"] = m
		w = data["weight"][masks["weight"]]
		w_s = data["weight_shape_sample"][masks["weight_shape_sample"]]
	axis_len = np.sqrt(np.sum(axis_v ** 2, axis=1))
	axis = (axis_v.transpose() / axis_len).transpose()  # :326-327
	if ellipticity == "distortion":
		e = (1 - q ** 2) / (1 + q ** 2)  # :357-358
	elif ellipticity == "ellipticity":
		e = (1 - q) / (1 + q)
	else:
		raise ValueError("Invalid value for ellipticity. Choose 'distortion' or 'ellipticity'.")
	return (np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray(pos_s, dtype=np.float64),
			np.ascontiguousarray(axis, dtype=np.float64), np.ascontiguousarray(e, dtype=np.float64),
			np.ascontiguousarray(w, dtype=np.float64), np.ascontiguousarray(w_s, dtype=np.float64))


def seq_sum(x):
	"""Python builtin ``sum`` over an array == strict left-to-right accumulation (measure_w_box_jk.py:364)."""
	x = np.asarray(x, dtype=np.float64)
	if x.size == 0:
		return 0
	return float(np.cumsum(x)[-1])  # numpy's cumsum is a sequential scan: same rounding as the builtin loop


def responsivity(w_s, e):
	"""R = sum(w (1 - e^2/2)) / sum(w): measure_w_box_jk.py:364."""
	return seq_sum(w_s * (1 - e ** 2 / 2.0)) / seq_sum(w_s)


def responsivity_jk(w_s, e, jk_s, num_box):
	"""measure_w_box_jk.py:463-466."""
	out = np.zeros(num_box)
	for i in range(num_box):
		m = np.where(jk_s != i)
		with np.errstate(invalid="ignore", divide="ignore"):
			out[i] = np.float64(seq_sum(w_s[m] * (1 - e[m] ** 2 / 2.0))) / np.float64(seq_sum(w_s[m]))
	return out


def jackknife_labels(pos, boxsize, L_subboxes):
	"""measure_IA_base.py:404-452: strict inequalities; points on any sub-box boundary keep label 0."""
	L_sub = (boxsize / 2.0) * 2.0 / L_subboxes
	lab = np.zeros(len(pos))
	num = 0
	for i in range(L_subboxes):
		for j in range(L_subboxes):
			for k in range(L_subboxes):
				xb = [i * L_sub, (i + 1) * L_sub]
				yb = [j * L_sub, (j + 1) * L_sub]
				zb = [k * L_sub, (k + 1) * L_sub]
				m = ((pos[:, 0] > xb[0]) * (pos[:, 0] < xb[1]) * (pos[:, 1] > yb[0]) * (pos[:, 1] < yb[1])
					 * (pos[:, 2] > zb[0]) * (pos[:, 2] < zb[1]))
				lab[m] = num
				num += 1
	return np.array(lab, dtype=int)


def random_pairs_rppi(r_bins, pi_bins, volume, n_pos, n_shape):
	"""get_random_pairs(..., 'cross', ...) on the whole grid: measure_IA_base.py:268-269, measure_w_box_jk.py:470-477."""
	n_r, n_p = len(r_bins) - 1, len(pi_bins) - 1
	rr = np.zeros((n_r, n_p))
	for i in range(n_r):
		for p in range(n_p):
			rr[i, p] = (n_pos * n_shape * np.pi * (r_bins[i + 1] ** 2 - r_bins[i] ** 2)
						* abs(pi_bins[p + 1] - pi_bins[p]) / volume)
	return rr


def _cap(mur, r):
	return np.pi / 3.0 * r ** 3 * (2 + mur) * (1 - mur) ** 2  # measure_IA_base.py:291


def random_pairs_rmu(r_bins, mu_bins, volume, n_pos, n_shape):
	"""get_random_pairs_r_mur(..., 'cross', ...): measure_IA_base.py:337-351 (note the (Np - 1) factor)."""
	n_r, n_m = len(r_bins) - 1, len(mu_bins) - 1
	rr = np.zeros((n_r, n_m))
	for i in range(n_r):
		for p in range(n_m):
			r_max, r_min, mur_max, mur_min = r_bins[i + 1], r_bins[i], mu_bins[p + 1], mu_bins[p]
			val = ((n_pos - 1.0) * n_shape
				   * (_cap(mur_min, r_max) - _cap(mur_max, r_max) - (_cap(mur_min, r_min) - _cap(mur_max, r_min)))
				   / volume)
			rr[i, p] = abs(val)
	return rr


def w_from_xi(xi, pi_bins):
	"""_measure_w_g_i: measure_IA_base.py:553-562."""
	dpi = pi_bins[1:] - pi_bins[:-1]
	dpi = np.array([dpi] * len(xi[:, 0]))
	return np.sum(xi * abs(dpi), axis=1)


def _assoc_legendre(l, m, x):
	"""P_l^m(x) for the two cases the reference uses ((2,2) and (0,0)); scipy's lpmn convention (Condon-Shortley)."""
	if (l, m) == (0, 0):
		return 1.0
	if (l, m) == (2, 2):
		return 3.0 * (1.0 - x * x)
	raise NotImplementedError


def multipoles_from_xi(xi, mu_r_bins, which):
	"""_measure_multipoles: measure_IA_base.py:628-655; which = 'g_plus' (l = s = 2) or 'gg' (l = s = 0)."""
	l = sab = 2 if which == "g_plus" else 0
	dmur = mu_r_bins[1:] - mu_r_bins[:-1]
	mu_mid = mu_r_bins[:-1] + abs(dmur / 2.0)  # the stored `_mu_r` bin centres, measure_m_box_jk.py:522-523
	n_r = xi.shape[0]
	Lg = np.zeros_like(xi)
	for n in range(n_r):
		for m in range(len(mu_mid)):
			Lg[n, m] = _assoc_legendre(l, sab, mu_mid[m])
	dmu = np.array(list(dmur) * n_r).reshape((n_r, len(dmur)))
	mult = (2 * l + 1) / 2.0 * math.factorial(l - sab) / math.factorial(l + sab) * Lg * xi * dmu
	return np.sum(mult, axis=1)


def combine_jackknife(realisations):
	"""_combine_jackknife_information: measure_IA_base.py:482-498.  realisations: [num_box, n_r]."""
	num_box, n_r = realisations.shape
	mean = np.zeros(n_r)
	for b in range(num_box):
		mean += realisations[b]
	mean /= num_box
	cov = np.zeros((n_r, n_r))
	std = np.zeros(n_r)
	for b in range(num_box):
		std += (realisations[b] - mean) ** 2
		for i in range(n_r):
			cov[:, i] += (realisations[b] - mean) * (realisations[b][i] - mean[i])
	std *= (num_box - 1) / num_box
	std = np.sqrt(std)
	cov *= (num_box - 1) / num_box
	return mean, std, cov


# --------------------------------------------------------------------------------------------------------------------
# the pair loop (C)
# --------------------------------------------------------------------------------------------------------------------
def _ptr(a, ct):
	return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def paircount(geom, pos, w, jk_p, pos_s, axis, e, w_s, jk_s, r_bins, sep_limits, bins2, boxsize, periodic, los,
			  two_R, num_box=0, rp_cut=0.0, s_range=None, use_grid=True, n_threads=1, r_thr=None, thr2=None):
	"""Five-tuple of the reference worker (measure_w_box_jk.py:646) + integer counts.

	Returns dict(DD, SpD, ScD, DD_jk, SpD_jk, count, n_tested, var, n_nan); SpD / ScD carry 1/(2R), SpD_jk does not;
	var = sum (w_D w_S e+ / 2R)^2, the brute variants' variance (measure_w_box_jk.py:196); n_nan = NaN-rule pairs.
	"""
	lib = _lib()
	n_r, n_2 = len(r_bins) - 1, len(bins2) - 1
	p = _Params(0 if geom == "rppi" else 1, n_r, n_2, float(sep_limits[0]), float(sep_limits[1]), float(r_bins[0]),
				float(r_bins[-1]), float(bins2[0]), float(bins2[-1]), float(boxsize), float(boxsize) / 2.0,
				1 if periodic else 0, int(los), float(rp_cut), int(num_box), float(two_R))
	nb = n_r * n_2
	njk = max(int(num_box), 0)
	DD, SpD, ScD = (np.zeros(nb) for _ in range(3))
	DD_jk, SpD_jk = np.zeros(max(njk, 1) * nb), np.zeros(max(njk, 1) * nb)
	count = np.zeros(nb, dtype=np.int64)
	var = np.zeros(nb)
	tested, n_nan = ctypes.c_int64(0), ctypes.c_int64(0)
	pos = np.ascontiguousarray(pos, dtype=np.float64)
	pos_s = np.ascontiguousarray(pos_s, dtype=np.float64)
	axis = np.ascontiguousarray(axis, dtype=np.float64)
	e = np.ascontiguousarray(e, dtype=np.float64)
	w = np.ascontiguousarray(w, dtype=np.float64)
	w_s = np.ascontiguousarray(w_s, dtype=np.float64)
	jp = np.ascontiguousarray(jk_p, dtype=np.int32) if jk_p is not None else None
	js = np.ascontiguousarray(jk_s, dtype=np.int32) if jk_s is not None else None
	s0, s1 = (0, len(pos_s)) if s_range is None else s_range
	rt = np.ascontiguousarray(r_thr, dtype=np.float64) if r_thr is not None else None
	t2 = np.ascontiguousarray(thr2, dtype=np.float64) if thr2 is not None else None
	D = ctypes.c_double
	rc = lib.oracle_paircount(ctypes.byref(p), ctypes.c_int64(len(pos)), _ptr(pos, D), _ptr(w, D),
							  _ptr(jp, ctypes.c_int32), ctypes.c_int64(len(pos_s)), _ptr(pos_s, D), _ptr(axis, D),
							  _ptr(e, D), _ptr(w_s, D), _ptr(js, ctypes.c_int32), ctypes.c_int64(s0),
							  ctypes.c_int64(s1), _ptr(rt, D), _ptr(t2, D), ctypes.c_int(1 if use_grid else 0),
							  ctypes.c_int(int(n_threads)), _ptr(DD, D), _ptr(SpD, D), _ptr(ScD, D), _ptr(DD_jk, D),
							  _ptr(SpD_jk, D), _ptr(count, ctypes.c_int64), ctypes.byref(tested), _ptr(var, D),
							  ctypes.byref(n_nan))
	if rc != 0:
		raise RuntimeError(f"oracle_paircount failed: {rc}")
	shp = (n_r, n_2)
	return dict(DD=DD.reshape(shp), SpD=SpD.reshape(shp), ScD=ScD.reshape(shp),
				DD_jk=DD_jk.reshape((max(njk, 1),) + shp)[:njk], SpD_jk=SpD_jk.reshape((max(njk, 1),) + shp)[:njk],
				count=count.reshape(shp), n_tested=int(tested.value), var=var.reshape(shp), n_nan=int(n_nan.value))


# --------------------------------------------------------------------------------------------------------------------
# full measurement: the same datasets the reference writes to HDF5, as {path: array}
# --------------------------------------------------------------------------------------------------------------------
def measure(data, kind, dataset_name="All", corr_type="both", num_jk=0, boxsize=None, snapshot=None,
			separation_limits=(0.1, 20.0), num_bins_r=8, num_bins_pi=20, pi_max=None, periodicity=True, masks=None,
			ellipticity="distortion", rp_cut=None, n_threads=1, use_grid=True, variant="tree"):
	"""Restatement of MeasureIABox.measure_xi_w / measure_xi_multipoles (measure_IA.py:68-262) with the *_tree
	variants underneath.  Returns {hdf5 path: ndarray} with the layout the reference writes."""
	data = dict(data)
	n_p, n_s = len(data["Position"]), len(data["Position_shape_sample"])
	data.setdefault("weight", np.ones(n_p))  # measure_IA_base.py:146-154
	data.setdefault("weight_shape_sample", np.ones(n_s))
	if corr_type not in ("both", "g+", "gg"):
		raise KeyError("Unknown value for corr_type. Choose from [g+, gg, both]")
	L_sub = 0
	if num_jk > 0:
		L_sub = round(num_jk ** (1.0 / 3))
		if L_sub ** 3 != num_jk:
			raise ValueError("Use x^3 as input for num_jk, with x as an int.")
	r_bins, pi_bins, mu_r_bins = make_bins(separation_limits, num_bins_r, num_bins_pi, pi_max, boxsize)
	pos, pos_s, axis, e, w, w_s = prepare(data, masks, ellipticity)
	los = int(data["LOS"])
	R = responsivity(w_s, e)
	geom = "rppi" if kind == "w" else "rmu"
	bins2 = pi_bins if geom == "rppi" else mu_r_bins
	jk_p = jk_s = None
	if num_jk > 0:
		jk_p = jackknife_labels(pos, boxsize, L_sub)
		jk_s = jackknife_labels(pos_s, boxsize, L_sub)
	res = paircount(geom, pos, w, jk_p, pos_s, axis, e, w_s, jk_s, r_bins, separation_limits, bins2, boxsize,
					periodicity, los, 2 * R, num_box=num_jk, rp_cut=0.0 if rp_cut is None else rp_cut,
					use_grid=use_grid, n_threads=n_threads)
	DD, SpD, ScD = res["DD"], res["SpD"], res["ScD"]
	L3 = boxsize ** 3
	Np, Ns = len(pos), len(pos_s)
	rr_fun = random_pairs_rppi if geom == "rppi" else random_pairs_rmu
	RR = rr_fun(r_bins, bins2, L3, Np, Ns)
	sep = r_bins[:-1] + abs((r_bins[1:] - r_bins[:-1]) / 2.0)
	mid2 = bins2[:-1] + abs((bins2[1:] - bins2[:-1]) / 2.0)
	top = "w" if geom == "rppi" else "multipoles"
	n1, n2 = ("_rp", "_pi") if geom == "rppi" else ("_r", "_mu_r")
	snap = f"Snapshot_{snapshot}/" if snapshot is not None else ""
	X = dataset_name
	out = {}
	zeros = np.zeros_like(DD)

	def put(group, name, arr):
		path = "/".join(p for p in (snap + group).split("/") if p) + "/" + name
		out[path] = np.array(arr, dtype=np.float64)

	jkg = f"{X}_jk{num_jk}" if num_jk > 0 else ""
	with np.errstate(divide="ignore", invalid="ignore"):
		put(f"{top}/xi_g_plus", X, SpD / RR)
		put(f"{top}/xi_g_plus", X + "_SplusD", SpD)
		put(f"{top}/xi_g_plus", X + "_RR_g_plus", RR)
		put(f"{top}/xi_g_plus", X + n1, sep)
		put(f"{top}/xi_g_plus", X + n2, mid2)
		put(f"{top}/xi_g_cross/{jkg}", X, ScD / RR)
		put(f"{top}/xi_g_cross/{jkg}", X + "_ScrossD", ScD)
		put(f"{top}/xi_g_cross/{jkg}", X + "_RR_g_cross", RR)
		put(f"{top}/xi_g_cross/{jkg}", X + n1, sep)
		put(f"{top}/xi_g_cross/{jkg}", X + n2, mid2)
		put(f"{top}/xi_gg", X, (DD / RR) - 1)
		put(f"{top}/xi_gg", X + "_DD", DD)
		put(f"{top}/xi_gg", X + "_RR_gg", RR)
		put(f"{top}/xi_gg", X + n1, sep)
		put(f"{top}/xi_gg", X + n2, mid2)
		if num_jk > 0:
			for grp in (f"{top}/xi_g_plus", f"{top}/xi_g_cross/{jkg}", f"{top}/xi_gg"):
				# tree variant: variance never accumulated (measure_w_box_jk.py:374,492); brute: sum term^2 / RR^2 (:196,:242)
				put(grp, X + "_sigmasq", res["var"] / RR ** 2 if variant == "brute" else zeros)
			R_jk = responsivity_jk(w_s, e, jk_s, num_jk)
			vol_jk = L3 * (num_jk - 1) / num_jk
			for i in range(num_jk):
				np_i = len(np.where(jk_p != i)[0])
				ns_i = len(np.where(jk_s != i)[0])
				RR_i = rr_fun(r_bins, bins2, vol_jk, np_i, ns_i)
				put(f"{top}/xi_g_plus/{jkg}", f"{X}_{i}", (SpD * (2 * R) - res["SpD_jk"][i]) / (RR_i * 2 * R_jk[i]))
				put(f"{top}/xi_g_plus/{jkg}", f"{X}_{i}_SplusD", (SpD * (2 * R) - res["SpD_jk"][i]) / (2 * R_jk[i]))
				put(f"{top}/xi_g_plus/{jkg}", f"{X}_{i}_RR", RR_i)
				put(f"{top}/xi_g_plus/{jkg}", f"{X}_{i}{n1}", sep)
				put(f"{top}/xi_g_plus/{jkg}", f"{X}_{i}{n2}", mid2)
				put(f"{top}/xi_gg/{jkg}", f"{X}_{i}", ((DD - res["DD_jk"][i]) / RR_i) - 1)
				put(f"{top}/xi_gg/{jkg}", f"{X}_{i}_DD", DD - res["DD_jk"][i])
				put(f"{top}/xi_gg/{jkg}", f"{X}_{i}_RR", RR_i)
				put(f"{top}/xi_gg/{jkg}", f"{X}_{i}{n1}", sep)
				put(f"{top}/xi_gg/{jkg}", f"{X}_{i}{n2}", mid2)

		# integrated statistics (measure_IA.py:136-149 / :234-247)
		kinds = {"both": ["g_plus", "gg"], "g+": ["g_plus"], "gg": ["gg"]}[corr_type]
		pre = "w_" if geom == "rppi" else "multipoles_"

		def integrate(xi, which):
			return w_from_xi(xi, pi_bins) if geom == "rppi" else multipoles_from_xi(xi, mu_r_bins, which)

		def key(group, name):
			return "/".join(p for p in (snap + group).split("/") if p) + "/" + name

		for which in kinds:
			xi = out[key(f"{top}/xi_{which}", X)]
			put(pre + which, X, integrate(xi, which))
			put(pre + which, X + n1, sep)
			if num_jk > 0:
				reals = []
				for i in range(num_jk):
					xi_i = out[key(f"{top}/xi_{which}/{jkg}", f"{X}_{i}")]
					val = integrate(xi_i, which)
					reals.append(val)
					put(f"{pre}{which}/{jkg}", f"{X}_{i}", val)
					put(f"{pre}{which}/{jkg}", f"{X}_{i}{n1}", sep)
				mean, std, cov = combine_jackknife(np.array(reals))
				put(pre + which, f"{X}_mean_{num_jk}", mean)
				put(pre + which, f"{X}_jackknife_{num_jk}", std)
				put(pre + which, f"{X}_jackknife_cov_{num_jk}", cov)
	out["__meta__/n_tested"] = np.array(res["n_tested"])
	out["__meta__/count"] = res["count"]
	return out

"""Seeded synthetic catalogues (SURVEY.md §8(d) "Concrete synthetic inputs").

The reference ships no input catalogue (its ``tests/conftest.py:19`` opens a file that is not in the repository), so
every parity test, golden fixture and bench line uses catalogues from here.  Only numpy; no GPU.
"""
from __future__ import annotations

import numpy as np


def uniform_box(n, boxsize, seed=1, n_shape=None, weights=False, los=2, clustered=0.0):
	"""Data dict in the reference's input format (reference ``measure_IA_base.py:81-95``).

	n, n_shape : sizes of the position and shape samples; ``n_shape=None`` means auto-correlation (the SAME array
	             object is used for both samples, as the reference's fixtures do).
	weights    : add ``weight`` / ``weight_shape_sample`` drawn from U(0.5, 1.5).
	clustered  : fraction of galaxies placed in Gaussian blobs (sigma = 2% of the box) instead of uniformly;
	             used for load-balance tests.
	"""
	rng = np.random.default_rng(seed)

	def positions(m):
		pos = rng.random((m, 3)) * boxsize
		if clustered > 0.0:
			k = int(m * clustered)
			centres = rng.random((max(1, k // 500), 3)) * boxsize
			which = rng.integers(0, len(centres), size=k)
			blob = centres[which] + rng.normal(scale=0.02 * boxsize, size=(k, 3))
			pos[:k] = np.mod(blob, boxsize)
		pos[pos >= boxsize] = np.nextafter(boxsize, 0.0)  # the reference's KDTree rejects x >= L
		return pos

	pos = positions(n)
	if n_shape is None:
		pos_s = pos
		ns = n
	else:
		pos_s = positions(n_shape)
		ns = n_shape
	theta = np.pi * rng.random(ns)
	length = rng.uniform(0.5, 2.0, ns)
	axis = np.stack([np.cos(theta), np.sin(theta)], axis=1) * length[:, None]  # un-normalised on purpose
	q = rng.uniform(0.2, 1.0, ns)
	data = {
		"Position": pos,
		"Position_shape_sample": pos_s,
		"Axis_Direction": axis,
		"LOS": int(los),
		"q": q,
	}
	if weights:
		data["weight"] = rng.uniform(0.5, 1.5, n)
		data["weight_shape_sample"] = data["weight"] if n_shape is None else rng.uniform(0.5, 1.5, ns)
	return data


def expected_pairs_rppi(n_pos, n_shape, boxsize, r_min, r_max, pi_lo, pi_hi):
	"""Expected number of binned ordered pairs for uniform randoms, (r_p, Pi) geometry."""
	return n_pos * n_shape * np.pi * (r_max ** 2 - r_min ** 2) * (pi_hi - pi_lo) / boxsize ** 3


def expected_pairs_rmu(n_pos, n_shape, boxsize, r_min, r_max):
	"""Expected number of binned ordered pairs for uniform randoms, (r, mu_r) geometry."""
	return n_pos * n_shape * 4.0 / 3.0 * np.pi * (r_max ** 3 - r_min ** 3) / boxsize ** 3

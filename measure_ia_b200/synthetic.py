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


def aligned_pairs_box(n, boxsize, seed=1, los=2):
	"""Catalogue that makes the reference's NaN rule fire (SURVEY.md section 8(a) hazard 3,
	``measure_w_box_jk.py:411-417``): galaxy 2k+1 sits a few Mpc from galaxy 2k, and the axis of galaxy 2k is its
	projected minimum-image separation from 2k+1 (times a random sign and length), i.e. exactly (anti)parallel.  The
	normalised dot product then exceeds 1 by an ulp for ~10 % of those pairs; the reference zeroes e+ / ex for them and
	still counts them in DD.  Auto-correlation (one array object for both samples)."""
	rng = np.random.default_rng(seed)
	n -= n % 2
	pos = rng.random((n, 3)) * boxsize
	off = rng.normal(size=(n // 2, 3))
	off *= (rng.uniform(0.3, 12.0, n // 2) / np.sqrt((off ** 2).sum(axis=1)))[:, None]
	pos[1::2] = np.mod(pos[0::2] + off, boxsize)
	pos[pos >= boxsize] = np.nextafter(boxsize, 0.0)
	not_los = [c for c in range(3) if c != los]
	theta = np.pi * rng.random(n)
	axis = np.stack([np.cos(theta), np.sin(theta)], axis=1) * rng.uniform(0.5, 2.0, n)[:, None]
	sep = pos[0::2][:, not_los] - pos[1::2][:, not_los]  # shape minus position, then the reference's two shifts
	sep[sep > boxsize / 2.0] -= boxsize
	sep[sep < -boxsize / 2.0] += boxsize
	axis[0::2] = sep * (rng.uniform(0.5, 2.0, n // 2) * rng.choice([-1.0, 1.0], n // 2))[:, None]
	return {"Position": pos, "Position_shape_sample": pos, "Axis_Direction": axis, "LOS": int(los),
			"q": rng.uniform(0.2, 1.0, n)}


def lattice_box(per_side, boxsize, seed=1, los=2, n_random=0):
	"""Every coordinate a multiple of boxsize / per_side (plus ``n_random`` uniform points): separations land EXACTLY
	on Pi bin edges, on +-L/2 (the periodic wrap's strict inequalities), on dz = 0 (mu_r = 0, an edge for an even
	number of mu_r bins), on r_p = r_min when separation_limits[0] is a lattice distance, and on jackknife faces
	(label-0 rule).  Auto-correlation."""
	rng = np.random.default_rng(seed)
	g = np.arange(per_side) * (boxsize / per_side)
	pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
	pos = pos[rng.permutation(len(pos))]
	if n_random:
		pos = np.concatenate([pos, rng.random((n_random, 3)) * boxsize])
	n = len(pos)
	theta = np.pi * rng.random(n)
	axis = np.stack([np.cos(theta), np.sin(theta)], axis=1) * rng.uniform(0.5, 2.0, n)[:, None]
	return {"Position": pos, "Position_shape_sample": pos, "Axis_Direction": axis, "LOS": int(los),
			"q": rng.uniform(0.2, 1.0, n)}


GENERATORS = {"uniform": uniform_box, "aligned_pairs": aligned_pairs_box, "lattice": lattice_box}

"""CPU: the three facts the symmetric auto-correlation kernels rest on (csrc/mia_tiled_rppi2s.cuh, mia_tiled_rmu.cuh SYM),
checked in numpy with the reference's operation sequence (measure_w_box_jk.py:401-407, measure_m_box_jk.py:418-431):

1. the wrapped separation of the reverse pair is EXACTLY the negated separation of the forward pair (so r_p^2, r^2 and the
   range tests are shared bit for bit, Pi -> -Pi, mu -> -mu);
2. the half-space rule "(d_u, d_v, d_z) lexicographically negative" takes every unordered pair with distinct positions
   exactly once;
3. with the calibrated thresholds, the bin of -x is the mirrored bin of x except ON an edge -- which is what the kernels'
   slow path handles (Pi: its own comparison against the mirrored edge; mu: the near-edge band)."""
import numpy as np

from measure_ia_b200 import calib


def _wrap(sep, L):
	sep = sep.copy()
	sep[sep > L / 2.0] -= L  # measure_w_box_jk.py:403
	sep[sep < -L / 2.0] += L  # :404
	return sep


def _catalogues():
	rng = np.random.default_rng(7)
	L = 50.0
	yield L, rng.random((400, 3)) * L
	# lattice: exact ties in every coordinate, separations of exactly +-L/2, coincident points
	g = np.arange(0, 8) * (L / 8.0)
	lat = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
	yield L, np.concatenate([lat[::3], lat[:5]])  # (with five duplicates)


def test_reverse_separation_is_the_exact_negative():
	for L, pos in _catalogues():
		fwd = _wrap(pos[:, None, :] - pos[None, :, :], L)  # sep[i, j] = pos_i - pos_j: shape i, position j
		rev = _wrap(pos[None, :, :] - pos[:, None, :], L)  # shape j, position i
		assert np.array_equal(rev, -fwd)
		rp2_f = fwd[..., 0] ** 2 + fwd[..., 1] ** 2
		rp2_r = rev[..., 0] ** 2 + rev[..., 1] ** 2
		assert np.array_equal(rp2_f, rp2_r)
		r2_f = (fwd ** 2).sum(axis=-1)
		assert np.array_equal(r2_f, (rev ** 2).sum(axis=-1))
		with np.errstate(invalid="ignore", divide="ignore"):
			mu_f, mu_r = fwd[..., 2] / np.sqrt(r2_f), rev[..., 2] / np.sqrt(r2_f)
		ok = r2_f > 0
		assert np.array_equal(mu_r[ok], -mu_f[ok])


def test_half_space_rule_takes_every_unordered_pair_once():
	for L, pos in _catalogues():
		d = _wrap(pos[:, None, :] - pos[None, :, :], L)
		du, dv, dz = d[..., 0], d[..., 1], d[..., 2]
		taken = (du < 0) | ((du == 0) & ((dv < 0) | ((dv == 0) & (dz < 0))))
		distinct = (du != 0) | (dv != 0) | (dz != 0)
		# exactly one of (i, j), (j, i) for distinct positions; none for coincident ones (r_p = 0 is never binned)
		assert np.array_equal(taken ^ taken.T, distinct)
		assert not (taken & taken.T).any()


def test_mirrored_bins_except_on_edges():
	rng = np.random.default_rng(11)
	for n in (8, 20, 5):
		# Pi bins: linspace(-pi_max, pi_max, n + 1) as in the reference (measure_IA_base.py), calibrated thresholds
		pi_bins = np.linspace(-102.5, 102.5, n + 1)
		thr = calib.pi_thresholds(pi_bins, n)
		x = np.concatenate([rng.uniform(-102.5, 102.5, 20000), pi_bins[1:-1], -pi_bins[1:-1], [0.0, -0.0]])

		def bin_of(v):  # number of interior thresholds passed (include/mia_b200.h), -1 / n outside the range
			b = (v[:, None] >= thr[None, 1:n]).sum(axis=1)
			return np.where((v >= thr[0]) & (v < thr[n]), b, -1)
		bf, br = bin_of(x), bin_of(-x)
		inside = (bf >= 0) & (br >= 0)
		on_edge = np.isin(np.abs(x), np.abs(pi_bins)) | (np.abs(np.abs(x)[:, None] - np.abs(thr)[None, :]).min(axis=1) < 1e-12)
		assert np.array_equal(br[inside & ~on_edge], n - 1 - bf[inside & ~on_edge])
		# and ON an edge the two orderings do NOT mirror (the reason the kernels compare the reverse pair on its own)
		k = n // 2
		if n % 2 == 0:
			e = np.array([pi_bins[k + 1]])
			assert bin_of(e)[0] + bin_of(-e)[0] != n - 1

"""CPU restatement of the (r, mu_r) kernel's division-free mu binning (measure_ia_b200/csrc/mia_tiled_rmu.cuh, rmu_approx):
t = fma(mu, n/2, n/2 + 6145) lies in [4096, 8192), where a double has exactly 40 fractional bits, so the bin is read from the
bit pattern and pairs within 16 * 2^-40 of a bin edge are flagged for the exact slow path.  The test checks the claim the kernel
relies on: whenever the pair is NOT flagged, the bit-pattern bin equals floor((mu + 1) n / 2) even if mu carries the 3e-16
error of the approximate reciprocal square root (reference formula: measure_m_box_jk.py:453-455)."""
import numpy as np
import pytest

BAND = np.uint32(16)


def bits_bin(mu, n_mu):
	hn = 0.5 * n_mu
	# fma emulated in extended precision (mu * hn is exact there: 53 + 6 bits), rounded once to double
	t = (mu.astype(np.longdouble) * np.longdouble(hn) + np.longdouble(hn + 6145.0)).astype(np.float64)
	b = t.view(np.uint64)
	thi = (b >> np.uint64(32)).astype(np.uint32)
	tlo = (b & np.uint64(0xFFFFFFFF)).astype(np.uint32)
	raw = (thi & np.uint32(0xFFFFF)) >> np.uint32(8)
	idx = np.minimum(np.maximum(raw, np.uint32(2049)), np.uint32(2048 + n_mu)).astype(np.int64) - 2049
	lo2 = (tlo + BAND).astype(np.uint32)
	h8 = (thi + (lo2 < BAND).astype(np.uint32)) & np.uint32(0xFF)
	return idx, (h8 == 0) & (lo2 < np.uint32(2) * BAND)


@pytest.mark.parametrize("n_mu", [1, 2, 5, 8, 10, 20])
def test_unflagged_pairs_get_the_exact_bin(n_mu):
	rng = np.random.default_rng(n_mu)
	edges = -1.0 + 2.0 * np.arange(n_mu + 1) / n_mu
	mu = np.concatenate([
		rng.uniform(-1, 1, 1_000_000),
		np.array([-1.0, 1.0, 0.0, -0.0, 1 - 1e-16, -1 + 1e-16, 1 + 2e-16, -1 - 2e-16]),
		edges, np.nextafter(edges, 5), np.nextafter(edges, -5),
		edges + 1e-12, edges - 1e-12, edges + 1e-10, edges - 1e-10, edges + 5e-12, edges - 5e-12])
	exact = np.clip(np.floor((mu.astype(np.longdouble) + 1) * np.longdouble(0.5 * n_mu)).astype(np.int64), 0, n_mu - 1)
	for err in (0.0, 3e-16, -3e-16):  # the kernel's mu is dz * rsqrt(s): relative error up to ~3e-16
		idx, flagged = bits_bin(mu * (1.0 + err), n_mu)
		assert idx.min() >= 0 and idx.max() <= n_mu - 1  # slot addresses stay inside the private histogram
		assert not ((idx != exact) & ~flagged).any()
		x = (mu.astype(np.longdouble) + 1) * np.longdouble(0.5 * n_mu)
		dist = np.abs(x - np.round(x)).astype(np.float64)
		assert flagged[dist < 5e-12].all()          # everything near an edge goes to the exact path ...
		assert not flagged[dist > 3e-11].any()      # ... and nothing else does (the slow path stays rare)
	assert bits_bin(rng.uniform(-1, 1, 1_000_000), n_mu)[1].mean() < 1e-9 * n_mu + 1e-6

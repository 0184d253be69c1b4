"""CPU: the reference's own golden HDF5 outputs (decoded to tests/golden/hdf5_*.npz) pin every host-side formula.

The input catalogue behind these files is not shipped with the reference (its tests/conftest.py:19 opens a file that
is absent), so pair counts cannot be regenerated; SURVEY.md section 4.1 lists the identities that must hold
bit-exactly inside the files.  They are checked here for the oracle's host functions and, in test_host_post.py, for
the product's.
"""
import numpy as np
import pytest

import parity_util as pu

CASES = [("mock_IA_TNG300", "All", 766), ("mock_IA_TNG300", "high", 292), ("mock_IA_TNG300", "low", 474),
		 ("mock_IA_TNG300_large", "All", 4547), ("mock_IA_TNG300_large", "high", 766),
		 ("mock_IA_TNG300_large", "low", 3781)]


@pytest.fixture(scope="module")
def bins(oracle):
	return oracle.make_bins((0.1, 20.0), 10, 8, None, 205.0)


@pytest.mark.parametrize("which,name,n", CASES)
def test_rr_xi_w(oracle, bins, which, name, n):
	g = pu.load_hdf5_fixture(which)
	r_bins, pi_bins, mu_bins = bins
	pre = "Snapshot_99/w/"
	rr = oracle.random_pairs_rppi(r_bins, pi_bins, 205.0 ** 3, n, n)
	assert np.array_equal(g[pre + f"xi_gg/{name}_RR_gg"], rr)
	assert np.array_equal(g[pre + f"xi_g_plus/{name}_RR_g_plus"], rr)
	assert np.array_equal(g[pre + f"xi_gg/{name}"], g[pre + f"xi_gg/{name}_DD"] / rr - 1)
	assert np.array_equal(g[pre + f"xi_g_plus/{name}"], g[pre + f"xi_g_plus/{name}_SplusD"] / rr)
	assert np.array_equal(g[f"Snapshot_99/w_gg/{name}"], oracle.w_from_xi(g[pre + f"xi_gg/{name}"], pi_bins))
	assert np.array_equal(g[f"Snapshot_99/w_g_plus/{name}"], oracle.w_from_xi(g[pre + f"xi_g_plus/{name}"], pi_bins))
	assert np.array_equal(g[f"Snapshot_99/w_gg/{name}_rp"], r_bins[:-1] + abs((r_bins[1:] - r_bins[:-1]) / 2.0))
	# DD is symmetric in Pi for an auto-correlation: every unordered pair is counted once per ordering
	dd = g[pre + f"xi_gg/{name}_DD"]
	assert np.array_equal(dd, dd[:, ::-1])
	assert dd.sum() == np.round(dd.sum())


@pytest.mark.parametrize("which,name,n", CASES)
def test_rr_xi_multipoles(oracle, bins, which, name, n):
	g = pu.load_hdf5_fixture(which)
	r_bins, pi_bins, mu_bins = bins
	pre = "Snapshot_99/multipoles/"
	rr = oracle.random_pairs_rmu(r_bins, mu_bins, 205.0 ** 3, n, n)
	assert np.array_equal(g[pre + f"xi_gg/{name}_RR_gg"], rr)
	assert np.array_equal(g[pre + f"xi_gg/{name}"], g[pre + f"xi_gg/{name}_DD"] / rr - 1)
	assert np.array_equal(g[f"Snapshot_99/multipoles_gg/{name}"],
						  oracle.multipoles_from_xi(g[pre + f"xi_gg/{name}"], mu_bins, "gg"))
	np.testing.assert_allclose(g[f"Snapshot_99/multipoles_g_plus/{name}"],
							   oracle.multipoles_from_xi(g[pre + f"xi_g_plus/{name}"], mu_bins, "g_plus"),
							   rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("which", ["mock_IA_TNG300", "mock_IA_TNG300_large"])
@pytest.mark.parametrize("stat", ["w_gg", "w_g_plus", "multipoles_gg", "multipoles_g_plus"])
def test_jackknife_combination(oracle, which, stat):
	"""_combine_jackknife_information recomputed from the stored realisations All_0..All_7 (old flat layout)."""
	g = pu.load_hdf5_fixture(which)
	reals = np.array([g[f"Snapshot_99/{stat}/All_{i}"] for i in range(8)])
	mean, std, cov = oracle.combine_jackknife(reals)
	assert np.array_equal(g[f"Snapshot_99/{stat}/All_mean_8"], mean)
	assert np.array_equal(g[f"Snapshot_99/{stat}/All_jackknife_cov_8"], cov)
	assert np.array_equal(g[f"Snapshot_99/{stat}/All_jackknife_8"], std)


def test_notebook_w_gg_vector():
	"""examples/example_MeasureIA_box.ipynb cell 13 prints w_gg for the 4547-galaxy mock."""
	g = pu.load_hdf5_fixture("mock_IA_TNG300_large")
	printed = [1216.05115556, 662.96525693, 486.19350429, 263.55473166, 138.65391819, 88.56665704, 56.12100847,
			   36.84167832, 19.65149456, 9.64465402]
	np.testing.assert_allclose(g["Snapshot_99/w_gg/All"], printed, rtol=1e-9)

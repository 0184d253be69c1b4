"""CPU: the oracle against the UNMODIFIED reference executed live (oracle/run_reference.py + oracle/ref_shims).

Runs only where /root/reference exists (the build container); skipped on the GPU box.  The committed fixtures
(tests/golden/ref_*.npz, tests/test_oracle_golden.py) pin the same thing everywhere else; this file re-derives a few of
them from the live reference so that a drifted fixture, shim or oracle shows up here first.
"""
import numpy as np
import pytest

import parity_util as pu
import run_reference

pytestmark = pytest.mark.skipif(not run_reference.reference_available(), reason="/root/reference is not present")


def _both(oracle, data, kind, boxsize, variant="tree", **kw):
	want = run_reference.run_reference(data, kind, boxsize=boxsize, variant=variant, **kw)
	got = oracle.measure(data, kind, boxsize=boxsize, variant=variant, n_threads=4, **kw)
	return got, want


@pytest.mark.parametrize("kind", ["w", "multipoles"])
@pytest.mark.parametrize("variant", ["tree", "brute"])
def test_live_reference_uniform(oracle, kind, variant):
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(1200, 120.0, seed=101)
	got, want = _both(oracle, data, kind, 120.0, variant=variant, num_jk=8, num_bins_r=6, num_bins_pi=6)
	pu.assert_datasets_match(got, want, exact_counts=True, label=f"live {kind}/{variant}: ")
	if variant == "brute":  # the variance output the tree variants leave at zero (measure_w_box_jk.py:196,242)
		top = "w" if kind == "w" else "multipoles"
		assert want[f"{top}/xi_g_plus/All_sigmasq"].max() > 0


@pytest.mark.parametrize("kind", ["w", "multipoles"])
def test_live_reference_nan_rule(oracle, kind):
	"""|c| > 1 by rounding zeroes e+ / ex and keeps the pair in DD (measure_w_box_jk.py:411-417,
	measure_m_box_jk.py:431-438): the oracle follows the live reference on a catalogue built to trip the rule."""
	from measure_ia_b200.synthetic import aligned_pairs_box
	data = aligned_pairs_box(1600, 80.0, seed=5)
	got, want = _both(oracle, data, kind, 80.0, num_jk=8, num_bins_r=6, num_bins_pi=6)
	pu.assert_datasets_match(got, want, exact_counts=True, label=f"live nan-rule {kind}: ")
	pos, pos_s, axis, e, w, w_s = oracle.prepare({**data, "weight": np.ones(1600), "weight_shape_sample": np.ones(1600)})
	r_bins, pi_bins, mu_bins = oracle.make_bins((0.1, 20.0), 6, 6, None, 80.0)
	res = oracle.paircount("rppi" if kind == "w" else "rmu", pos, w, None, pos_s, axis, e, w_s, None, r_bins, (0.1, 20.0),
						   pi_bins if kind == "w" else mu_bins, 80.0, True, 2, 1.0)
	assert res["n_nan"] >= 50, res["n_nan"]


def test_live_reference_masks_weights_cross(oracle):
	"""BASELINE.json configs[4] in small: two different samples, weights, masks, LOS = 1."""
	import make_golden
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(1500, 100.0, seed=102, n_shape=900, weights=True, los=1)
	masks = make_golden.make_masks(data, 7)
	got, want = _both(oracle, data, "w", 100.0, num_jk=8, num_bins_r=5, num_bins_pi=6, masks=masks)
	pu.assert_datasets_match(got, want, exact_counts=False, label="live masks: ")


@pytest.mark.parametrize("kind,estimator,corr", [("w", "galaxies", "both"), ("multipoles", "clusters", "both"), ("w", "clusters", "gg")])
def test_live_reference_lightcone(tmp_path, monkeypatch, kind, estimator, corr):
	"""Light-cone brute loops: oracle/pylightcone.py + the product's host side against the live reference on a
	catalogue that is NOT one of the committed fixtures (pyccl replaced by oracle/ref_shims/pyccl on both sides)."""
	import make_golden_lightcone as mg
	from measure_ia_b200.lightcone import MeasureIALightcone
	from test_host_post import read_all
	from test_lightcone_host import compare_with_fixture, oracle_lc_pair_sums, run_product
	cat = dict(n=350, n_shape=280, n_rand=500, seed=901, weights=True)
	call = dict(kind=kind, IA_estimator=estimator, corr_type=corr, over_h=(kind == "multipoles"))
	want = mg.run_reference(cat, call)
	monkeypatch.setattr(MeasureIALightcone, "_pair_sums", oracle_lc_pair_sums())
	out = str(tmp_path / "lc.hdf5")
	run_product(dict(catalogue=cat, call=call, binning=mg.BINNING), out)
	compare_with_fixture(read_all(out), want, f"live light-cone {kind}/{estimator}/{corr}: ", False)

"""CPU: the oracle (oracle/pyoracle.py + oracle.c) against outputs of the unmodified reference (tests/golden/ref_*)."""
import numpy as np
import pytest

import parity_util as pu


@pytest.mark.parametrize("name", pu.fixture_names())
def test_oracle_matches_reference_fixture(oracle, name):
	meta, want = pu.load_fixture(name)
	data, masks, kw = pu.rebuild_inputs(meta)
	kind = kw.pop("kind")
	variant = kw.pop("variant", "tree")  # brute: `_sigmasq` = sum term^2 / RR^2 (measure_w_box_jk.py:196,242); tree: zeros
	unit_weights = "weight" not in data
	got = oracle.measure(data, kind, boxsize=meta["catalogue"]["boxsize"], masks=masks, n_threads=4, variant=variant, **kw)
	pu.assert_datasets_match(got, want, exact_counts=unit_weights, label=f"{name}: ")


def test_oracle_brute_equals_grid(oracle):
	"""Reference tests/test_w_sim_internal_consistency.py:59-79 (brute == tree): candidate search must not matter."""
	from measure_ia_b200.synthetic import uniform_box
	d = uniform_box(1200, 100.0, seed=21)
	for kind in ("w", "multipoles"):
		a = oracle.measure(d, kind, num_jk=8, boxsize=100.0, num_bins_r=6, num_bins_pi=6, use_grid=True)
		b = oracle.measure(d, kind, num_jk=8, boxsize=100.0, num_bins_r=6, num_bins_pi=6, use_grid=False)
		assert np.array_equal(a["__meta__/count"], b["__meta__/count"])
		assert b["__meta__/n_tested"] == 1200 * 1200
		pu.assert_datasets_match(a, {k: v for k, v in b.items() if not k.startswith("__meta__")})


def test_jackknife_labels_known_answer(oracle):
	"""Reference tests/test_w_jk.py:16-23 with tests/conftest.py:59-69: box 3, 2^3 regions -> [0, 5, 7, 3]."""
	com = np.array([[1, 1, 1], [2, 1, 2], [2.5, 2.5, 1.51], [1, 2, 2]], dtype=float)
	assert list(oracle.jackknife_labels(com, 3.0, 2)) == [0, 5, 7, 3]

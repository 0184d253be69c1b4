"""Helpers shared by the parity tests: golden fixture loading and the comparison rule.

Comparison rule (BASELINE.json north_star; SURVEY.md section 8(a) hazard 4):
  * pair counts (``*_DD`` with unit weights, integer ``count``) must be bit-exact;
  * every other dataset: |a - b| <= rtol * |a| + atol_scale * max|a|, rtol = 1e-10.  The absolute term is tied to the
    largest entry of the array because S+D / SxD bins are sums of random-sign terms and may cancel to ~0.
"""
import json
import os
import sys

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(_REPO, "tests", "golden")
sys.path.insert(0, os.path.join(_REPO, "oracle"))

RTOL = 1e-10
ATOL_SCALE = 1e-11


def load_fixture(name):
	z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
	meta = json.loads(str(z["__meta__"]))
	out = {k.replace("|", "/"): z[k] for k in z.files if k != "__meta__"}
	return meta, out


def load_hdf5_fixture(which):
	z = np.load(os.path.join(GOLDEN, f"hdf5_{which}.npz"))
	return {k.replace("|", "/"): z[k] for k in z.files}


def fixture_names():
	return sorted(f[4:-4] for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))


def rebuild_inputs(meta):
	"""Regenerate the seeded catalogue (and masks) a fixture was made from and verify its digest."""
	import make_golden
	data, masks, kw = make_golden.build_inputs(meta["catalogue"], meta["measurement"])
	assert make_golden.input_digest(data, masks) == meta["digest"], "synthetic generator drifted from the fixture"
	return data, masks, kw


def assert_datasets_match(got, want, exact_counts=True, rtol=RTOL, atol_scale=ATOL_SCALE, label=""):
	"""got / want: {hdf5 path: array}.  Every dataset of `want` must exist in `got` and agree."""
	missing = sorted(set(want) - set(got))
	assert not missing, f"{label}: datasets missing: {missing[:8]} (+{max(0, len(missing) - 8)})"
	for k in sorted(want):
		a, b = np.asarray(want[k], dtype=np.float64), np.asarray(got[k], dtype=np.float64)
		assert a.shape == b.shape, f"{label}{k}: shape {b.shape} != {a.shape}"
		with np.errstate(all="ignore"):
			assert np.array_equal(np.isnan(a), np.isnan(b)), f"{label}{k}: NaN pattern differs"
			inf = np.isinf(a)
			assert np.array_equal(inf, np.isinf(b)) and np.array_equal(a[inf], b[inf]), f"{label}{k}: inf differs"
			m = np.isfinite(a)
			if not m.any():
				continue
			if exact_counts and k.endswith("_DD"):
				assert np.array_equal(a, b), f"{label}{k}: pair counts differ (max |d| = {np.abs(a - b).max()})"
				continue
			scale = np.abs(a[m]).max()
			tol = rtol * np.abs(a[m]) + atol_scale * scale
			err = np.abs(a[m] - b[m])
			worst = np.argmax(err - tol)
			assert (err <= tol).all(), (f"{label}{k}: |d|={err[worst]:.3e} > tol={tol[worst]:.3e} "
										 f"(value {a[m][worst]:.6e}, array scale {scale:.3e})")

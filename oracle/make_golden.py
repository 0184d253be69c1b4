"""Generate tests/golden/*.npz.  Run in the build container (needs /root/reference); the fixtures are committed.

Two kinds of fixture:
  ref_<name>.npz      outputs of the UNMODIFIED reference (oracle/run_reference.py) on seeded synthetic catalogues
                      (measure_ia_b200.synthetic.uniform_box).  Stored: the config (json), a sha256 of the inputs, and
                      every dataset the reference wrote to its HDF5 file.
  hdf5_<file>.npz     the reference's own golden HDF5 outputs (tests/data/processed/TNG300/*.hdf5) decoded with
                      measure_ia_b200.h5lite; they pin the analytic-RR / xi / w / multipole / jackknife formulas.

Usage:  python oracle/make_golden.py [name ...]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _REPO)
sys.path.insert(0, _HERE)

from measure_ia_b200.synthetic import GENERATORS, uniform_box  # noqa: E402
import run_reference  # noqa: E402

GOLDEN = os.path.join(_REPO, "tests", "golden")

# name -> (catalogue kwargs, measurement kwargs)
CONFIGS = {
	"w_auto_jk27": (dict(n=3000, boxsize=205.0, seed=1), dict(kind="w", num_jk=27, num_bins_r=10, num_bins_pi=8)),
	"m_auto_jk27": (dict(n=3000, boxsize=205.0, seed=1), dict(kind="multipoles", num_jk=27, num_bins_r=10, num_bins_pi=8)),
	"w_auto_nojk_default_bins": (dict(n=3000, boxsize=205.0, seed=2), dict(kind="w", num_jk=0)),
	"m_auto_nojk_default_bins": (dict(n=3000, boxsize=205.0, seed=2), dict(kind="multipoles", num_jk=0)),
	"w_cross_weights_los0_jk8": (dict(n=2000, n_shape=1500, boxsize=205.0, seed=3, weights=True, los=0),
								 dict(kind="w", num_jk=8, num_bins_r=10, num_bins_pi=8)),
	"m_cross_weights_los1_jk8": (dict(n=2000, n_shape=1500, boxsize=205.0, seed=4, weights=True, los=1),
								 dict(kind="multipoles", num_jk=8, num_bins_r=10, num_bins_pi=8)),
	"w_pimax30_jk8": (dict(n=3000, boxsize=205.0, seed=5), dict(kind="w", num_jk=8, num_bins_r=6, num_bins_pi=6, pi_max=30.0)),
	"w_nonperiodic": (dict(n=2500, boxsize=100.0, seed=6), dict(kind="w", num_jk=8, num_bins_r=5, num_bins_pi=4,
																 periodicity=False, separation_limits=(0.5, 15.0))),
	"m_nonperiodic": (dict(n=2500, boxsize=100.0, seed=6), dict(kind="multipoles", num_jk=8, num_bins_r=5, num_bins_pi=4,
																 periodicity=False, separation_limits=(0.5, 15.0))),
	"w_ellipticity_def": (dict(n=2000, boxsize=75.0, seed=7), dict(kind="w", num_jk=27, num_bins_r=6, num_bins_pi=10,
																	ellipticity="ellipticity", separation_limits=(0.2, 10.0))),
	"w_masked": (dict(n=3000, boxsize=205.0, seed=8, weights=True), dict(kind="w", num_jk=8, num_bins_r=10, num_bins_pi=8,
																		 mask_seed=11)),
	"m_masked": (dict(n=3000, boxsize=205.0, seed=8, weights=True), dict(kind="multipoles", num_jk=8, num_bins_r=10,
																		 num_bins_pi=8, mask_seed=11)),
	"w_small_box": (dict(n=400, boxsize=30.0, seed=9), dict(kind="w", num_jk=8, num_bins_r=4, num_bins_pi=4,
															 separation_limits=(0.5, 12.0), variant="brute")),
	"m_small_box": (dict(n=400, boxsize=30.0, seed=9), dict(kind="multipoles", num_jk=8, num_bins_r=4, num_bins_pi=4,
															 separation_limits=(0.5, 12.0), variant="brute")),
	"w_auto_jk27_30k": (dict(n=30000, boxsize=205.0, seed=1), dict(kind="w", num_jk=27, num_bins_r=10, num_bins_pi=8)),
	"m_auto_jk27_30k": (dict(n=30000, boxsize=205.0, seed=1), dict(kind="multipoles", num_jk=27, num_bins_r=10, num_bins_pi=8)),
	# the reference's NaN rule (|c| > 1 by rounding -> e+ = ex = 0, pair still counted; measure_w_box_jk.py:411-417,
	# measure_m_box_jk.py:431-438): ~1500 exactly (anti)parallel pairs, of which > 100 trip it
	"w_nan_rule": (dict(gen="aligned_pairs", n=3000, boxsize=100.0, seed=12),
				   dict(kind="w", num_jk=8, num_bins_r=10, num_bins_pi=8)),
	"m_nan_rule": (dict(gen="aligned_pairs", n=3000, boxsize=100.0, seed=12),
				   dict(kind="multipoles", num_jk=8, num_bins_r=10, num_bins_pi=8)),
	"w_nan_rule_los0": (dict(gen="aligned_pairs", n=2000, boxsize=60.0, seed=13, los=0),
						dict(kind="w", num_jk=27, num_bins_r=6, num_bins_pi=6)),
	# exact edges: lattice coordinates put Pi on bin edges and on +-L/2, dz = 0 (mu_r on the central edge), r_p = r_min,
	# points on jackknife faces; the brute variant keeps r_p == r_min (the tree variant drops it, SURVEY.md 8(a) hazard 7)
	"w_lattice": (dict(gen="lattice", per_side=16, boxsize=40.0, seed=14, n_random=300),
				  dict(kind="w", num_jk=8, num_bins_r=5, num_bins_pi=8, separation_limits=(2.5, 15.0), variant="brute")),
	"m_lattice": (dict(gen="lattice", per_side=16, boxsize=40.0, seed=14, n_random=300),
				  dict(kind="multipoles", num_jk=8, num_bins_r=5, num_bins_pi=8, separation_limits=(2.5, 15.0),
					   variant="brute")),
	"w_lattice_los1_pimax": (dict(gen="lattice", per_side=12, boxsize=30.0, seed=15, los=1, n_random=150),
							 dict(kind="w", num_jk=27, num_bins_r=4, num_bins_pi=6, separation_limits=(2.5, 12.5), pi_max=7.5,
								  variant="brute")),
}


def make_masks(data, mask_seed):
	"""Boolean masks in the form of reference tests/test_masks.py:15-18 (same key set as the data dict)."""
	rng = np.random.default_rng(mask_seed)
	n_p, n_s = len(data["Position"]), len(data["Position_shape_sample"])
	same = data["Position"] is data["Position_shape_sample"]
	mp = rng.random(n_p) < 0.7
	ms = mp if same else rng.random(n_s) < 0.6
	return {"Position": mp, "Position_shape_sample": ms, "Axis_Direction": ms, "q": ms, "weight": mp,
			"weight_shape_sample": ms}


def build_inputs(cat_kw, meas_kw):
	cat_kw = dict(cat_kw)
	gen = GENERATORS[cat_kw.pop("gen", "uniform")]
	first = cat_kw.pop("n") if "n" in cat_kw else cat_kw.pop("per_side")
	data = gen(first, cat_kw.pop("boxsize"), **cat_kw)
	meas_kw = dict(meas_kw)
	masks = None
	if "mask_seed" in meas_kw:
		masks = make_masks(data, meas_kw.pop("mask_seed"))
	return data, masks, meas_kw


def input_digest(data, masks):
	h = hashlib.sha256()
	for k in sorted(data):
		h.update(k.encode())
		h.update(np.ascontiguousarray(data[k]).tobytes())
	if masks:
		for k in sorted(masks):
			h.update(np.ascontiguousarray(masks[k]).tobytes())
	return h.hexdigest()


def generate(name):
	cat_kw, meas_kw = CONFIGS[name]
	data, masks, kw = build_inputs(cat_kw, meas_kw)
	kind = kw.pop("kind")
	t = time.time()
	out = run_reference.run_reference(data, kind, boxsize=cat_kw["boxsize"], masks=masks, **kw)
	meta = dict(name=name, catalogue=cat_kw, measurement=meas_kw, digest=input_digest(data, masks),
				numpy=np.__version__, seconds=round(time.time() - t, 2))
	path = os.path.join(GOLDEN, f"ref_{name}.npz")
	np.savez_compressed(path, __meta__=np.array(json.dumps(meta)), **{k.replace("/", "|"): v for k, v in out.items()})
	print(f"{name}: {len(out)} datasets, {meta['seconds']} s -> {os.path.relpath(path, _REPO)} "
		  f"({os.path.getsize(path) // 1024} KiB)")


def generate_covariance_fixture():
	"""BASELINE.json config 5 post-step: three projections of one box (dataset names LOS_x / LOS_y / LOS_z), jackknife
	realisations by the reference's measure_xi_w / measure_xi_multipoles, then the unmodified
	MeasureJackknife.create_full_cov_matrix_projections / measure_covariance_multiple_datasets on the same file."""
	import tempfile
	from measure_ia_b200 import h5lite
	measureia = run_reference.load_reference()
	tmp = tempfile.mkdtemp(prefix="mia_cov_")
	out = os.path.join(tmp, "cov.hdf5")
	devnull, stdout = open(os.devnull, "w"), sys.stdout
	sys.stdout = devnull
	try:
		for los, name in enumerate(("LOS_x", "LOS_y", "LOS_z")):
			data = uniform_box(1500, 100.0, seed=40 + los, los=los)
			box = measureia.MeasureIABox(data, out, None, 7, [0.5, 15.0], 5, 6, None, 100.0, True, 1)
			box.measure_xi_w(name, "both", num_jk=8, temp_file_path=tmp + "/")
			box.measure_xi_multipoles(name, "both", num_jk=8, temp_file_path=tmp + "/")
		jk = measureia.MeasureJackknife(None, out, None, 7, [0.5, 15.0], 5, 6, None, 100.0)
		for corr in ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg"):
			jk.create_full_cov_matrix_projections(corr, ["LOS_x", "LOS_y", "LOS_z"], num_box=8)
	finally:
		sys.stdout = stdout
		devnull.close()
	# the reference never closes the handle it opened in create_full_cov_matrix_projections (:603); flush it
	import gc
	gc.collect()
	f = h5lite.File(out, "r")
	flat = run_reference._flatten(f)
	f.close()
	keep = {k: v for k, v in flat.items() if k.split("/")[1] in ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg")}
	path = os.path.join(GOLDEN, "cov_projections.npz")
	np.savez_compressed(path, **{k.replace("/", "|"): v for k, v in keep.items()})
	print(f"covariance fixture: {len(keep)} datasets -> {os.path.relpath(path, _REPO)} ({os.path.getsize(path) // 1024} KiB)")


def decode_reference_hdf5():
	from measure_ia_b200 import h5lite
	src = "/root/reference/tests/data/processed/TNG300"
	for fn in ("mock_IA_TNG300.hdf5", "mock_IA_TNG300_large.hdf5"):
		f = h5lite.File(os.path.join(src, fn), "r")
		flat = run_reference._flatten(f)
		f.close()
		path = os.path.join(GOLDEN, "hdf5_" + fn.replace(".hdf5", ".npz"))
		np.savez_compressed(path, **{k.replace("/", "|"): v for k, v in flat.items()})
		print(f"{fn}: {len(flat)} datasets -> {os.path.relpath(path, _REPO)} ({os.path.getsize(path) // 1024} KiB)")


if __name__ == "__main__":
	os.makedirs(GOLDEN, exist_ok=True)
	names = sys.argv[1:] or list(CONFIGS) + ["hdf5", "cov"]
	for n in names:
		if n == "hdf5":
			decode_reference_hdf5()
		elif n == "cov":
			generate_covariance_fixture()
		else:
			generate(n)

"""BASELINE.json configs[4] end to end on the GPU: a shape sample x density sample CROSS-correlation with weights and masks,
measured along the three lines of sight as the datasets LOS_x / LOS_y / LOS_z (w and multipoles, 27 jackknife regions), then
``create_full_cov_matrix_projections`` (reference measure_jackknife.py:573-648) on the file the GPU path wrote.

Checked against the CPU oracle (oracle/pyoracle.py) on the same seeded inputs: every dataset of every projection (pair
counts of the weighted catalogue to 1e-10, like all sums), and the combined 3-projection covariance matrices against the
same combination applied to the oracle's jackknife realisations."""
import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu

N_POS, N_SHAPE, L, NUM_JK = 200_000, 50_000, 205.0, 27
NAMES = ["LOS_x", "LOS_y", "LOS_z"]


def _inputs():
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(N_POS, L, seed=505, n_shape=N_SHAPE, weights=True)
	rng = np.random.default_rng(506)
	masks = {"Position": rng.random(N_POS) < 0.7, "Position_shape_sample": rng.random(N_SHAPE) < 0.6}
	masks["Axis_Direction"] = masks["q"] = masks["Position_shape_sample"]
	masks["weight"] = masks["Position"]
	masks["weight_shape_sample"] = masks["Position_shape_sample"]
	return data, masks


def _read_all(path):
	from measure_ia_b200 import h5lite
	f = h5lite.File(path, "r")
	out = {}

	def walk(g, pre):
		for k, v in g.items():
			if isinstance(v, h5lite.Group):
				walk(v, pre + k + "/")
			else:
				out[pre + k] = v[...]
	walk(f, "")
	f.close()
	return out


def test_cfg5_cross_weights_masks_three_projections(tmp_path, oracle):
	import torch
	if not torch.cuda.is_available():
		pytest.skip("GPU tests need a CUDA device")
	from measure_ia_b200 import MeasureIABox, MeasureJackknife, h5lite
	data, masks = _inputs()
	out = str(tmp_path / "cfg5.hdf5")
	ref_path = str(tmp_path / "cfg5_oracle.hdf5")
	fref, written = h5lite.File(ref_path, "w"), set()
	box = MeasureIABox(data, out, boxsize=L, num_bins_r=10, num_bins_pi=8)
	for los, name in enumerate(NAMES):
		data["LOS"] = los
		for kind, run in (("w", box.measure_xi_w), ("multipoles", box.measure_xi_multipoles)):
			run(name, "both", num_jk=NUM_JK, temp_file_path=False, masks=dict(masks))
			assert box.last_stats["kernel"] == 2, "cross-correlations take the ordered tiled kernels"
			want = oracle.measure(data, kind, dataset_name=name, num_jk=NUM_JK, boxsize=L, num_bins_r=10, num_bins_pi=8,
								  masks=dict(masks), n_threads=oracle.max_threads(), variant="brute")
			count = want.pop("__meta__/count")
			want.pop("__meta__/n_tested")
			assert np.array_equal(box.last_result["count"], count), f"{name}/{kind}: pair counts differ"
			pu.assert_datasets_match(_read_all(out), want, exact_counts=False, label=f"cfg5 {name}/{kind}: ")
			for k, v in want.items():  # the oracle's realisations, for the combination step below
				if f"_jk{NUM_JK}/" in k or k.endswith(f"{name}_jackknife_cov_{NUM_JK}"):
					if k not in written:
						written.add(k)
						fref.create_dataset(k, data=v)
	fref.close()
	# ---- full jackknife covariance of the three projections (measure_jackknife.py:573-648) -----------------------------
	for corr in ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg"):
		box.create_full_cov_matrix_projections(corr, NAMES, num_box=NUM_JK)
		MeasureJackknife(None, ref_path, None, None, [0.1, 20.0], 10, 8, None, L).create_full_cov_matrix_projections(
			corr, NAMES, num_box=NUM_JK)
	got, want = _read_all(out), _read_all(ref_path)
	combined = {k: v for k, v in want.items() if "combined_jackknife_cov" in k or k.split("/")[-1].count("LOS_") == 2}
	assert len(combined) == 4 * (4 + 3 * 2), sorted(combined)  # per statistic: 4 block matrices + 3 pair cov + 3 pair std
	pu.assert_datasets_match(got, combined, exact_counts=False, label="cfg5 combined covariance: ")
	cov3 = got[f"w_g_plus/LOS_x_LOS_y_LOS_z_combined_jackknife_cov_{NUM_JK}"]
	assert cov3.shape == (30, 30)


def test_batched_projections_equal_separate_calls(tmp_path):
	"""``measure_xi_projections``: one catalogue preparation for three projections x two statistics, one file
	handle, the three-projection covariance at the end -- the file must hold what the reference's workflow (one run per
	projection and statistic with that projection's shapes, then ``create_full_cov_matrix_projections``) writes."""
	import torch
	if not torch.cuda.is_available():
		pytest.skip("GPU tests need a CUDA device")
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	n_pos, n_shape = 60_000, 25_000
	data = uniform_box(n_pos, L, seed=707, n_shape=n_shape, weights=True)
	rng = np.random.default_rng(708)
	proj = []
	for los in range(3):  # projected shapes differ from one line of sight to the next
		th = np.pi * rng.random(n_shape)
		proj.append({"LOS": los, "Axis_Direction": np.stack([np.cos(th), np.sin(th)], 1) * rng.uniform(0.5, 2.0, n_shape)[:, None],
					 "q": rng.uniform(0.2, 1.0, n_shape)})
	masks = {"Position": rng.random(n_pos) < 0.7, "Position_shape_sample": rng.random(n_shape) < 0.6}
	masks["Axis_Direction"] = masks["q"] = masks["weight_shape_sample"] = masks["Position_shape_sample"]
	masks["weight"] = masks["Position"]
	kw = dict(boxsize=L, num_bins_r=10, num_bins_pi=8)

	sep = str(tmp_path / "separate.hdf5")
	counts = {}
	for name, p in zip(NAMES, proj):
		d = dict(data)
		d.update(p)
		box = MeasureIABox(d, sep, **kw)
		for kind, run in (("w", box.measure_xi_w), ("multipoles", box.measure_xi_multipoles)):
			run(name, "both", num_jk=NUM_JK, temp_file_path=False, masks=dict(masks))
			counts[(name, kind)] = (box.last_result["count"].copy(), box.last_result["count_jk"].copy())
	for corr in ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg"):
		box.create_full_cov_matrix_projections(corr, NAMES, num_box=NUM_JK)

	bat = str(tmp_path / "batched.hdf5")
	d = dict(data)
	d["LOS"] = 2  # not used: every projection names its own line of sight
	box = MeasureIABox(d, bat, **kw)
	box.measure_xi_projections(NAMES, "both", num_jk=NUM_JK, temp_file_path=False, masks=dict(masks), projections=proj)
	assert set(box.last_results) == set(counts)
	for key, (c, cjk) in counts.items():
		assert np.array_equal(box.last_results[key]["count"], c) and np.array_equal(box.last_results[key]["count_jk"], cjk), key
		assert box.last_stats["per_measurement"][key]["kernel"] == 2
	got, want = _read_all(bat), _read_all(sep)
	assert set(got) == set(want)
	pu.assert_datasets_match(got, want, exact_counts=False, label="batched projections: ")

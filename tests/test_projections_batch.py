"""CPU: host logic of ``MeasureIABox.measure_xi_projections`` (several projections of one box in one call).

The operator is replaced by the CPU oracle (test-only stand-ins for the three device-side hooks), so what is compared is
everything the batched call adds AROUND the pair loop: one catalogue preparation shared by all projections, per-projection
shapes (``Axis_Direction`` / ``q`` / ``LOS``), one file handle for all datasets, and the three-projection covariance
(reference workflow: three ``MeasureIABox`` runs, then ``MeasureJackknife.create_full_cov_matrix_projections``,
measure_jackknife.py:573-648).  The file must hold exactly what the separate calls write.  The GPU version of this test
is tests/test_gpu_cfg5.py::test_batched_projections_equal_separate_calls."""
import numpy as np
import pytest

from measure_ia_b200 import MeasureIABox, h5lite
from test_host_post import oracle_pair_sums, read_all

NAMES = ["LOS_x", "LOS_y", "LOS_z"]
L, NUM_JK = 60.0, 8


def _catalogue(seed=41, n=900, n_shape=400):
	from measure_ia_b200.synthetic import uniform_box
	d = uniform_box(n, L, seed=seed, n_shape=n_shape, weights=True)
	rng = np.random.default_rng(seed + 1)
	# projected shapes differ from one line of sight to the next, as in the reference's catalogues
	proj = []
	for los in range(3):
		th = np.pi * rng.random(n_shape)
		proj.append({"LOS": los, "Axis_Direction": np.stack([np.cos(th), np.sin(th)], 1) * rng.uniform(0.5, 2.0, n_shape)[:, None],
					 "q": rng.uniform(0.2, 1.0, n_shape)})
	masks = {"Position": rng.random(n) < 0.8, "Position_shape_sample": rng.random(n_shape) < 0.7}
	masks["Axis_Direction"] = masks["q"] = masks["weight_shape_sample"] = masks["Position_shape_sample"]
	masks["weight"] = masks["Position"]
	return d, proj, masks


def _install_host_hooks(monkeypatch, oracle, counters):
	"""Host (numpy + oracle) versions of the three hooks the batched call uses on the device."""
	single = oracle_pair_sums(oracle, n_threads=1)  # fixed summation order: the two files must agree bit for bit

	def _device(self):
		return "cpu"

	def _prepare_catalogue(self, masks, L_subboxes, dev):
		counters["catalogue"] += 1
		return dict(masks=masks, L=L_subboxes)

	def _prepare_shapes(self, C, axis_direction, q_ratio, masks, ellipticity):
		counters["shapes"] += 1
		return dict(C, axis_direction=axis_direction, q=q_ratio)

	def _pair_sums(self, geom, masks, L_subboxes, ellipticity, rp_cut=None, variance=False, prepared=None, los=None):
		if prepared is None:
			return single(self, geom, masks, L_subboxes, ellipticity, rp_cut, variance)
		counters["pairs"] += 1
		saved = {k: self.data[k] for k in ("Axis_Direction", "q", "LOS")}
		self.data.update(Axis_Direction=prepared["axis_direction"], q=prepared["q"], LOS=los)
		try:
			return single(self, geom, masks, L_subboxes, ellipticity, rp_cut, variance)
		finally:
			self.data.update(saved)

	for name, fn in (("_device", _device), ("_prepare_catalogue", _prepare_catalogue), ("_prepare_shapes", _prepare_shapes),
					 ("_pair_sums", _pair_sums)):
		monkeypatch.setattr(MeasureIABox, name, fn)


@pytest.mark.parametrize("use_masks", [False, True])
def test_batched_call_writes_what_separate_calls_write(oracle, tmp_path, monkeypatch, use_masks):
	counters = dict(catalogue=0, shapes=0, pairs=0)
	_install_host_hooks(monkeypatch, oracle, counters)
	d, proj, masks = _catalogue()
	masks = masks if use_masks else None
	kw = dict(boxsize=L, num_bins_r=5, num_bins_pi=6, separation_limits=[0.5, 12.0])

	# ---- the reference's workflow: one run per projection and statistic, then the combination ----------------------------
	sep = str(tmp_path / "separate.hdf5")
	for name, p in zip(NAMES, proj):
		dd = dict(d)
		dd.update(p)
		box = MeasureIABox(dd, sep, **kw)
		box.measure_xi_w(name, "both", num_jk=NUM_JK, temp_file_path=False, masks=None if masks is None else dict(masks))
		box.measure_xi_multipoles(name, "both", num_jk=NUM_JK, temp_file_path=False, masks=None if masks is None else dict(masks))
	for corr in ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg"):
		box.create_full_cov_matrix_projections(corr, NAMES, num_box=NUM_JK)

	# ---- the batched call -------------------------------------------------------------------------------------------------
	bat = str(tmp_path / "batched.hdf5")
	dd = dict(d)
	dd["LOS"] = 2  # ignored by the batched call: every projection names its own line of sight
	box = MeasureIABox(dd, bat, **kw)
	box.measure_xi_projections(NAMES, "both", num_jk=NUM_JK, temp_file_path=False, masks=None if masks is None else dict(masks),
							   projections=proj)
	assert counters == dict(catalogue=1, shapes=3, pairs=6)  # ONE catalogue preparation for six measurements
	assert dd["LOS"] == 2 and set(box.last_results) == {(n, s) for n in NAMES for s in ("w", "multipoles")}

	got, want = read_all(bat), read_all(sep)
	assert set(got) == set(want)
	for k in want:
		a, b = got[k], want[k]
		assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)), k
		assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), k
	assert got[f"w_g_plus/LOS_x_LOS_y_LOS_z_combined_jackknife_cov_{NUM_JK}"].shape == (15, 15)


def test_batched_call_options_and_errors(oracle, tmp_path, monkeypatch):
	counters = dict(catalogue=0, shapes=0, pairs=0)
	_install_host_hooks(monkeypatch, oracle, counters)
	d, proj, _ = _catalogue(seed=7, n=300, n_shape=300)
	out = str(tmp_path / "o.hdf5")
	box = MeasureIABox(dict(d), out, boxsize=L, num_bins_r=4, num_bins_pi=4, separation_limits=[0.5, 10.0])
	# two projections, one statistic, no jackknife: no combination step, default lines of sight 0, 1
	box.measure_xi_projections(["a", "b"], "g+", num_jk=0, temp_file_path=False, statistics="w")
	keys = read_all(out)
	assert "w_g_plus/a" in keys and "w_g_plus/b" in keys and not any("combined" in k or k.startswith("multipoles") for k in keys)
	assert counters == dict(catalogue=1, shapes=2, pairs=2)
	with pytest.raises(ValueError, match="temp_file_path"):
		box.measure_xi_projections(NAMES, "both", num_jk=8)
	with pytest.raises(ValueError, match="x\\^3"):
		box.measure_xi_projections(NAMES, "both", num_jk=10, temp_file_path=False)
	with pytest.raises(KeyError):
		box.measure_xi_projections(NAMES, "g++", temp_file_path=False)
	with pytest.raises(KeyError):
		box.measure_xi_projections(NAMES, "both", temp_file_path=False, statistics=["xi"])
	with pytest.raises(ValueError, match="projections"):
		box.measure_xi_projections(NAMES, "both", temp_file_path=False, projections=proj[:2])
	with pytest.raises(ValueError, match="LOS"):
		box.measure_xi_projections(["a"], "both", temp_file_path=False, projections=[{"LOS": 3}])


def test_batched_call_needs_a_gpu(tmp_path):
	"""No CPU fallback: without the stand-ins the call fails loudly on a CUDA-less machine."""
	import torch
	if torch.cuda.is_available():
		pytest.skip("CUDA device present")
	d, proj, _ = _catalogue(seed=3, n=50, n_shape=50)
	box = MeasureIABox(dict(d), str(tmp_path / "x.hdf5"), boxsize=L)
	with pytest.raises(RuntimeError, match="CUDA"):
		box.measure_xi_projections(NAMES, "both", temp_file_path=False, projections=proj)

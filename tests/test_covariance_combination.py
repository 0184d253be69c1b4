"""CPU: multi-dataset jackknife covariance (measure_ia_b200/jackknife.py) against the unmodified reference
(tests/golden/cov_projections.npz, made by oracle/make_golden.py from MeasureJackknife on three projections) and against
the known answers documented in the reference's (commented-out) tests/test_covariance_multiple_datasets.py:15-38."""
import numpy as np
import pytest

from measure_ia_b200 import MeasureIABox, MeasureJackknife, h5lite

CORRS = ("w_g_plus", "w_gg", "multipoles_g_plus", "multipoles_gg")


def _fixture():
	z = np.load(__file__.rsplit("/", 1)[0] + "/golden/cov_projections.npz")
	return {k.replace("|", "/"): z[k] for k in z.files}


def test_projection_covariances_match_reference(tmp_path):
	want = _fixture()
	path = str(tmp_path / "cov.hdf5")
	f = h5lite.File(path, "w")
	own = tuple(f"LOS_{a}_jackknife_cov_8" for a in "xyz")
	for k, v in want.items():  # the realisations and each dataset's own covariance (written by measure_xi_w) go in;
		if "_jk8/" in k or k.endswith(own):  # every combined product must be regenerated
			f.create_dataset(k, data=v)
	f.close()
	jk = MeasureJackknife(None, path, None, 7, [0.5, 15.0], 5, 6, None, 100.0)
	for corr in CORRS:
		jk.create_full_cov_matrix_projections(corr, ["LOS_x", "LOS_y", "LOS_z"], num_box=8)
	f = h5lite.File(path, "r")
	checked = 0
	for k, v in want.items():
		if "_jk8/" in k or not ("jackknife_cov_8" in k or k.endswith("_jackknife_8")):
			continue
		parts = k.split("/")
		name = parts[-1]
		if name in ("LOS_x_jackknife_cov_8", "LOS_y_jackknife_cov_8", "LOS_z_jackknife_cov_8", "LOS_x_jackknife_8",
					"LOS_y_jackknife_8", "LOS_z_jackknife_8"):
			continue  # written by measure_xi_w itself (covered by the ref_* fixtures), not by the combination
		got = f[k][:]
		assert got.shape == v.shape, k
		assert np.array_equal(np.isnan(got), np.isnan(v)), k
		m = ~np.isnan(v)
		assert np.array_equal(got[m], v[m]), k
		checked += 1
	f.close()
	assert checked == 4 * (3 * 2 + 4)  # per statistic: 3 pair cov + 3 pair std + 4 combined matrices


def test_known_answers_three_realisations(tmp_path):
	"""Toy data of the reference's commented-out test: cov(set1) = 2/3 [[2,1,1,-1],[1,2,2,-2],[1,2,2,-2],[-1,-2,-2,2]]."""
	a = np.array([[1.0, 2, 3, 4], [2.0, 4, 5, 2], [3.0, 3, 4, 3]])  # realisation b, bin i
	path = str(tmp_path / "toy.hdf5")
	f = h5lite.File(path, "w")
	for name, arr in (("set1", a), ("set2", a + 10.0), ("set3", 2 * a)):
		g = f.create_group(f"Snapshot_99/w_g_plus/{name}_jk3")
		for b in range(3):
			g.create_dataset(f"{name}_{b}", data=arr[b])
	f.close()
	box = MeasureIABox(None, path, "TNG100", 99, [0.1, 20.0], 4, 8)
	cov1, std1 = box.measure_covariance_multiple_datasets("w_g_plus", ["set1"], 3, True)
	cov2, _ = box.measure_covariance_multiple_datasets("w_g_plus", ["set2"], 3, True)
	cov3, _ = box.measure_covariance_multiple_datasets("w_g_plus", ["set3"], 3, True)
	cov13, _ = box.measure_covariance_multiple_datasets("w_g_plus", ["set1", "set3"], 3, True)
	d = a - a.mean(0)
	np.testing.assert_allclose(cov1, 2 / 3 * d.T @ d, rtol=1e-14, atol=1e-15)
	np.testing.assert_allclose(cov2, cov1, rtol=1e-12, atol=1e-14)
	np.testing.assert_allclose(cov3, 4 * cov1, rtol=1e-14, atol=1e-15)
	np.testing.assert_allclose(cov13, 2 * cov1, rtol=1e-14, atol=1e-15)
	np.testing.assert_allclose(std1, np.sqrt(np.diag(cov1)), rtol=1e-14)
	with pytest.raises(ValueError):
		box.measure_covariance_multiple_datasets("w_gx", ["set1"], 3, True)
	with pytest.raises(KeyError):
		box.measure_covariance_multiple_datasets("w_gg", ["a", "b", "c"], 3, True)

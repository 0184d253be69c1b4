"""CPU: the product's host side (measure_ia_b200/box.py, io.py, h5lite.py, calib.py) without any GPU work.

The pair sums are injected from the oracle (test-only) so that everything AROUND the operator -- input preparation,
jackknife labels, responsivities, analytic randoms, xi / w / multipoles / covariance and the HDF5 layout -- is compared
with the unmodified reference's output files (tests/golden/ref_*.npz)."""
import ctypes
import os

import numpy as np
import pytest

import parity_util as pu
from measure_ia_b200 import MeasureIABox, SimInfo, h5lite
from measure_ia_b200.box import integer_cube_root


def oracle_pair_sums(oracle, n_threads=4):
	"""A stand-in for MeasureIABox._pair_sums that gets the five accumulators from the CPU oracle (n_threads=1: sums in a
	fixed order, bit-reproducible from call to call)."""
	def _pair_sums(self, geom, masks, L_subboxes, ellipticity, rp_cut=None, variance=False):
		pos, pos_s, axis, e, w, w_s, same = self._prepare(masks, ellipticity)
		num_box = L_subboxes ** 3 if L_subboxes else 0
		jk_p = jk_s = None
		if num_box:
			jk_p = self._jackknife_labels(pos, L_subboxes)
			jk_s = self._jackknife_labels(pos_s, L_subboxes)
		R, R_jk = self._responsivity(w_s, e, jk_s, num_box)
		bins2 = self.pi_bins if geom == "rppi" else self.mu_r_bins
		r = oracle.paircount(geom, pos, w, jk_p, pos_s, axis, e, w_s, jk_s, self.r_bins, (self.r_min, self.r_max), bins2,
							 self.boxsize, self.periodicity, int(self.data["LOS"]), 1.0, num_box=num_box, n_threads=n_threads)
		self.last_stats = dict(rank=0)
		return dict(count=r["count"], DD=r["DD"], SpD_raw=r["SpD"], ScD_raw=r["ScD"], count_jk=None, DD_jk=r["DD_jk"],
					SpD_jk=r["SpD_jk"], var_raw=r["var"] if variance else None, R=R, R_jk=R_jk, jk_p=jk_p, jk_s=jk_s,
					Np=len(pos), Ns=len(pos_s))
	return _pair_sums


def read_all(path):
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


@pytest.mark.parametrize("name", [n for n in pu.fixture_names() if "30k" not in n])
def test_host_pipeline_reproduces_reference_files(oracle, tmp_path, monkeypatch, name):
	meta, want = pu.load_fixture(name)
	data, masks, kw = pu.rebuild_inputs(meta)
	kind = kw.pop("kind")
	variant = kw.pop("variant", "tree")
	num_jk = kw.pop("num_jk", 0)
	ellipticity = kw.pop("ellipticity", "distortion")
	out = str(tmp_path / "out.hdf5")
	monkeypatch.setattr(MeasureIABox, "_pair_sums", oracle_pair_sums(oracle))
	box = MeasureIABox(data, out, None, None, list(kw.pop("separation_limits", (0.1, 20.0))), kw.pop("num_bins_r", 8),
					   kw.pop("num_bins_pi", 20), kw.pop("pi_max", None), meta["catalogue"]["boxsize"],
					   kw.pop("periodicity", True))
	assert not kw, kw
	run = box.measure_xi_w if kind == "w" else box.measure_xi_multipoles
	# brute variants: temp_file_path=False, `_sigmasq` = sum (w_D w_S e+ / 2R)^2 / RR^2; tree variants: zeros
	run("All", "both", num_jk=num_jk, temp_file_path=False if variant == "brute" else str(tmp_path) + "/", masks=masks,
		ellipticity=ellipticity)
	got = read_all(out)
	assert set(got) == set(want) | {k for k in got if k.endswith("_sigmasq")} or set(want) <= set(got)
	pu.assert_datasets_match(got, want, exact_counts="weight" not in meta["catalogue"] and not meta["catalogue"].get("weights"),
							 label=f"{name}: ")


def test_jackknife_labels(oracle):
	"""Known answer of reference tests/test_w_jk.py:16-23 and equality with the n^3 strict-inequality loops, including
	points exactly on sub-box faces (label 0)."""
	com = np.array([[1, 1, 1], [2, 1, 2], [2.5, 2.5, 1.51], [1, 2, 2]], dtype=float)
	d = {"Position": com, "Position_shape_sample": com, "Axis_Direction": np.array([]), "LOS": 2, "q": np.array([])}
	box = MeasureIABox(d, None, None, None, boxsize=3.0)
	a, b = box._get_jackknife_region_indices(None, 2)
	assert list(a) == [0, 5, 7, 3] and list(b) == [0, 5, 7, 3]
	rng = np.random.default_rng(5)
	pts = rng.random((5000, 3)) * 90.0
	pts[:50, 0] = 30.0  # on a face
	pts[50:80, 2] = 0.0
	pts[80:100, 1] = 60.0
	pts[100:110] = [30.0, 60.0, 0.0]
	d = {"Position": pts, "Position_shape_sample": pts[:100], "Axis_Direction": np.zeros((100, 2)), "LOS": 2,
		 "q": np.ones(100)}
	box = MeasureIABox(d, None, None, None, boxsize=90.0)
	for n in (1, 2, 3, 4):
		a, b = box._get_jackknife_region_indices(None, n)
		assert np.array_equal(a, oracle.jackknife_labels(pts, 90.0, n))
		assert np.array_equal(b, oracle.jackknife_labels(pts[:100], 90.0, n))


def test_error_behaviour(tmp_path):
	from measure_ia_b200.synthetic import uniform_box
	d = uniform_box(50, 100.0, seed=1)
	box = MeasureIABox(d, str(tmp_path / "x.hdf5"), boxsize=100.0)
	with pytest.raises(ValueError, match="x\\^3"):
		box.measure_xi_w("a", "both", num_jk=10, temp_file_path=False)
	with pytest.raises(ValueError, match="temp_file_path"):
		box.measure_xi_w("a", "both", num_jk=8)  # temp_file_path=None (measure_IA.py:108-110)
	with pytest.raises(ValueError, match="temp_file_path"):
		box.measure_xi_multipoles("a", "both")
	with pytest.raises(KeyError):
		box.measure_xi_w("a", "g++", num_jk=0, temp_file_path=False)
	with pytest.raises(ValueError, match="pi_max and boxsize"):
		MeasureIABox(d, None)
	assert integer_cube_root(27) == (3, True) and integer_cube_root(64) == (4, True) and not integer_cube_root(9)[1]


def test_default_weights_injected_into_callers_dict():
	from measure_ia_b200.synthetic import uniform_box
	d = uniform_box(10, 50.0, seed=1)
	assert "weight" not in d
	MeasureIABox(d, None, boxsize=50.0)
	assert np.array_equal(d["weight"], np.ones(10)) and np.array_equal(d["weight_shape_sample"], np.ones(10))


def test_sim_info_matrix():
	"""Reference tests/test_sim_input.py:4-46."""
	s = SimInfo("TNG300", 99)
	assert (s.simname, s.snapshot, s.boxsize, s.L_0p5, s.snap_group) == ("TNG300", "99", 205.0, 102.5, "Snapshot_99/")
	s = SimInfo("TNG100", None)
	assert (s.boxsize, s.snap_group, s.snapshot) == (75.0, "", None)
	s = SimInfo(None, 40, boxsize=100.0)
	assert (s.simname, s.boxsize, s.L_0p5, s.snap_group) == (None, 100.0, 50.0, "Snapshot_40/")
	s = SimInfo(None, None)
	assert s.boxsize is None and s.L_0p5 is None
	assert SimInfo("EAGLE", 28).boxsize == 100.0 * 0.6777
	assert SimInfo("FLAMINGO_L1", 1).boxsize == 1000.0 * 0.681
	with pytest.raises(KeyError):
		SimInfo("Illustris", 1)


def test_golden_identities_with_product_formulas():
	g = pu.load_hdf5_fixture("mock_IA_TNG300")
	box = MeasureIABox(None, None, "TNG300", 99, [0.1, 20], 10, 8)
	assert np.array_equal(g["Snapshot_99/w/xi_gg/All_RR_gg"], box._rr_grid_rppi(205.0 ** 3, 766, 766))
	assert np.array_equal(g["Snapshot_99/multipoles/xi_gg/All_RR_gg"], box._rr_grid_rmu(205.0 ** 3, 766, 766))
	assert np.array_equal(g["Snapshot_99/w_gg/All"], box._w_from_xi(g["Snapshot_99/w/xi_gg/All"], box.pi_bins))
	assert np.array_equal(g["Snapshot_99/multipoles_gg/All"],
						  box._multipole_from_xi(g["Snapshot_99/multipoles/xi_gg/All"], box.mu_r_bins, "gg"))
	np.testing.assert_allclose(g["Snapshot_99/multipoles_g_plus/All"],
							   box._multipole_from_xi(g["Snapshot_99/multipoles/xi_g_plus/All"], box.mu_r_bins, "g_plus"),
							   rtol=1e-13, atol=1e-16)
	reals = np.array([g[f"Snapshot_99/w_g_plus/All_{i}"] for i in range(8)])
	mean, std, cov = box._jackknife_stats(reals)
	np.testing.assert_allclose(g["Snapshot_99/w_g_plus/All_jackknife_cov_8"], cov, rtol=1e-13, atol=0)
	np.testing.assert_allclose(g["Snapshot_99/w_g_plus/All_jackknife_8"], std, rtol=1e-13)


def test_combine_jackknife_from_file(tmp_path):
	"""Reference tests/test_w_jk.py:5-13: _combine_jackknife_information on realisations stored in the current layout."""
	g = pu.load_hdf5_fixture("mock_IA_TNG300")
	path = str(tmp_path / "jk.hdf5")
	f = h5lite.File(path, "w")
	grp = f.create_group("Snapshot_99/w_g_plus/All_jk8")
	for i in range(8):
		grp.create_dataset(f"All_{i}", data=g[f"Snapshot_99/w_g_plus/All_{i}"])
	f.close()
	box = MeasureIABox(None, path, "TNG300", 99, [0.1, 20], 10, 8)
	covs, stds = box._combine_jackknife_information("All", "All_jk8", ["w_g_plus"], 8, return_output=True)
	np.testing.assert_allclose(covs[0], g["Snapshot_99/w_g_plus/All_jackknife_cov_8"], rtol=1e-13)
	box._combine_jackknife_information("All", "All_jk8", ["w_g_plus"], 8)
	f = h5lite.File(path, "r")
	np.testing.assert_allclose(f["Snapshot_99/w_g_plus/All_mean_8"][:], g["Snapshot_99/w_g_plus/All_mean_8"], rtol=1e-15)
	f.close()


def test_h5lite_overwrite_and_many_members(tmp_path):
	path = str(tmp_path / "t.hdf5")
	f = h5lite.File(path, "a")
	g = f.create_group("a/b")
	for i in range(700):  # forces several symbol-table nodes and a two-level B-tree
		g.create_dataset(f"d_{i}", data=np.arange(i % 7 + 1, dtype=np.float64) * i)
	f.close()
	f = h5lite.File(path, "a")
	assert len(f["a/b"]) == 700 and np.array_equal(f["a//b/d_13"][:], np.arange(7.0) * 13)
	del f["a/b"]["d_13"]
	f["a/b"].create_dataset("d_13", data=np.ones((3, 2)))
	f["a"].create_dataset("ints", data=np.arange(5, dtype=np.int32))
	f.close()
	f = h5lite.File(path, "r")
	assert f["a/b/d_13"].shape == (3, 2) and f["a/ints"].dtype == np.int32 and len(f["a/b"]) == 700
	f.close()


def test_library_exports_every_declared_symbol():
	"""The C-ABI library loads without a GPU and exports every function include/mia_b200.h declares."""
	import re
	from measure_ia_b200 import ops
	from measure_ia_b200.build import build_library, HEADER
	lib = ctypes.CDLL(build_library())
	header = open(HEADER).read()
	declared = set(re.findall(r"\b(mia_[a-z0-9_]+)\s*\(", header))
	assert declared == set(ops.EXPORTS), declared ^ set(ops.EXPORTS)
	for sym in declared:
		assert hasattr(lib, sym), sym
	assert ops.load_library().mia_abi_version() == ops.MIA_ABI_VERSION
	assert ops.load_library().mia_strerror(-3).decode().startswith("a coordinate")


def test_operator_refuses_cpu_tensors():
	import torch
	from measure_ia_b200 import ops  # noqa: F401  (registers the op)
	z = torch.zeros((4, 3), dtype=torch.float64)
	thr = torch.tensor([0.0, 1.0, 2.0], dtype=torch.float64)
	with pytest.raises(RuntimeError, match="CUDA|cuda|No CUDA|device"):
		torch.ops.measure_ia_b200.paircount(z, None, None, z, None, None, z[:, :2].contiguous(), z[:, 0].contiguous(), thr,
											thr, 0, 2, True, 0, 10.0, 2.0, 0.0, 0, 0, 1)


def test_workspace_planning_runs_without_a_gpu(monkeypatch):
	"""mia_workspace_bytes plans the grid and the kernel choice on the host only (no compute call): every configuration the
	tiled kernels take or decline must yield a workspace size, forced-tiled requests on declined configurations must not,
	and the size must grow with the catalogue."""
	import torch
	from measure_ia_b200 import MeasureIABox, ops
	from measure_ia_b200.synthetic import uniform_box
	lib = ops.load_library()
	box = MeasureIABox(uniform_box(10, 205.0, seed=1), None, boxsize=205.0, num_bins_r=10, num_bins_pi=8)

	def ws(geom, n, num_jk=27, kernel="auto", n_2=8):
		b = box if n_2 == 8 else MeasureIABox(uniform_box(10, 205.0, seed=1), None, boxsize=205.0, num_bins_r=10, num_bins_pi=n_2)
		r2_thr, thr2, rp2_cut, _ = b._thresholds_for(geom, None)
		p = ops.make_params(ops.GEOM_RPPI if geom == "rppi" else ops.GEOM_RMU, 10, n_2, 2, True, num_jk, ops.KERNEL_NAMES[kernel],
							205.0, float(b.r_bins[-1]), float(rp2_cut), torch.from_numpy(r2_thr), torch.from_numpy(thr2))
		return lib.mia_workspace_bytes(ctypes.byref(p), n, n)

	import ctypes
	for geom in ("rppi", "rmu"):
		sizes = [ws(geom, n) for n in (1000, 100000, 1000000)]
		assert all(s > 0 for s in sizes) and sizes[0] < sizes[1] < sizes[2], (geom, sizes)
		assert ws(geom, 100000, num_jk=0) > 0 and ws(geom, 100000, kernel="general") > 0
	# both (r_p, Pi) kernels plan; 40 mu bins exceed the tiled (r, mu_r) kernel's private slots: auto falls back, forced fails
	for mode in ("0", "2"):
		monkeypatch.setenv("MIA_RPPI_V2", mode)
		assert ws("rppi", 1000000, kernel="tiled") > 0
	assert ws("rmu", 100000, n_2=40) > 0
	assert ws("rmu", 100000, n_2=40, kernel="tiled") == 0
	# planning knobs that change the workspace: the symmetric kernels' 64-byte candidate records (equal sample sizes only),
	# the accumulator cap, finer worker slots for large catalogues
	monkeypatch.delenv("MIA_RPPI_V2")
	full = ws("rppi", 1000000)
	assert ws("rppi", 1000000, kernel="tiled_ordered") < full  # 32-byte candidates, no symmetric path
	monkeypatch.setenv("MIA_SLOT_MULT", "1")
	one = ws("rppi", 1000000)
	assert one < full and full - one > 1.0e9  # 4x the accumulator copies by default at 1e6 shapes (1.96 GB vs 0.49 GB)
	monkeypatch.delenv("MIA_SLOT_MULT")
	monkeypatch.setenv("MIA_ACC_CAP_MB", "256")
	assert ws("rppi", 1000000) < one
	monkeypatch.delenv("MIA_ACC_CAP_MB")
	assert ws("rppi", 1000000, num_jk=1000) < 6.0e9  # 2000 rows x 80 bins x 32 B per copy: capped at 4 GiB of copies


def test_h5lite_reuses_the_tree_it_wrote_last(tmp_path):
	"""Re-opening a file this process wrote and nobody touched since skips the parse; changes made through the new handle
	are written, and a file replaced on disk is parsed again."""
	import os
	import time
	from measure_ia_b200 import h5lite
	p = os.path.join(str(tmp_path), "a.hdf5")
	f = h5lite.File(p, "a")
	f.create_group("w/xi").create_dataset("A", data=np.arange(6.).reshape(2, 3))
	f.close()
	f = h5lite.File(p, "a")  # cache hit: modify a nested and a top-level member, then read back with a FRESH parse
	assert "w/xi/A" in f
	del f["w/xi/A"]
	f["w/xi"].create_dataset("A", data=np.ones(4))
	f.create_dataset("top", data=np.arange(3))
	f.close()
	h5lite._RECENT.clear()
	f = h5lite.File(p, "r")
	assert np.array_equal(f["w/xi/A"][:], np.ones(4)) and np.array_equal(f["top"][:], np.arange(3))
	f.close()
	f = h5lite.File(p, "a")
	f.create_dataset("x", data=np.zeros(2))
	f.close()
	time.sleep(0.01)  # another writer replaces the file: the cached tree must not be used
	q = os.path.join(str(tmp_path), "b.hdf5")
	f2 = h5lite.File(q, "w")
	f2.create_dataset("other", data=np.arange(5))
	f2.close()
	os.replace(q, p)
	f = h5lite.File(p, "r")
	assert list(f.keys()) == ["other"]
	f.close()
	f = h5lite.File(p, "a")
	f.create_dataset("y", data=np.arange(2))
	f.close()
	f = h5lite.File(p, "r")  # read-only reopen right after a write: served from the cache
	assert sorted(f.keys()) == ["other", "y"]
	f.close()

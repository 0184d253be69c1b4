"""GPU parity tests proper: the CUDA path (through the torch op and through the raw C ABI) against
(a) outputs of the unmodified reference (tests/golden/ref_*.npz), (b) the CPU oracle on the same seeded inputs, and
(c) size-independent properties at larger sizes."""
import ctypes
import os

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu

KERNELS = ["general", "auto", "tiled_ordered"]
TILED = (2, 4)  # reported kernel ids: tiled, tiled with the symmetric auto-correlation path (mia_b200.h)


@pytest.fixture(scope="module")
def torch_cuda():
	import torch
	if not torch.cuda.is_available():
		pytest.skip("GPU tests need a CUDA device (run with -m gpu on the B200 box)")
	from measure_ia_b200 import ops
	ops.load_library()
	return torch


def read_all(path):
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


def run_box(meta, data, masks, kw, out, kernel):
	from measure_ia_b200 import MeasureIABox
	kw = dict(kw)
	kind = kw.pop("kind")
	# the reference's brute variants (temp_file_path=False) also accumulate the `_sigmasq` variance; a temporary path selects
	# the tree variants, which store zeros there (measure_IA.py:102-131)
	temp = False if kw.pop("variant", None) == "brute" else os.path.dirname(out) + "/"
	num_jk = kw.pop("num_jk", 0)
	ellipticity = kw.pop("ellipticity", "distortion")
	box = MeasureIABox(data, out, None, None, list(kw.pop("separation_limits", (0.1, 20.0))), kw.pop("num_bins_r", 8),
					   kw.pop("num_bins_pi", 20), kw.pop("pi_max", None), meta["catalogue"]["boxsize"],
					   kw.pop("periodicity", True))
	assert not kw, kw
	box.kernel = kernel
	run = box.measure_xi_w if kind == "w" else box.measure_xi_multipoles
	run("All", "both", num_jk=num_jk, temp_file_path=temp, masks=masks, ellipticity=ellipticity)
	return box


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", pu.fixture_names())
def test_reference_fixture(torch_cuda, tmp_path, name, kernel):
	"""Whole API path on the GPU vs the files the unmodified reference wrote for the same inputs."""
	meta, want = pu.load_fixture(name)
	data, masks, kw = pu.rebuild_inputs(meta)
	out = str(tmp_path / "out.hdf5")
	box = run_box(meta, data, masks, kw, out, kernel)
	got = read_all(out)
	pu.assert_datasets_match(got, want, exact_counts=not meta["catalogue"].get("weights"), label=f"{name}[{kernel}]: ")
	if kernel != "general":
		assert box.last_stats["kernel"] in TILED, "the tiled kernels should cover every (r_p, Pi) and (r, mu_r) fixture"
		assert kernel == "auto" or box.last_stats["kernel"] == 2
		if kw.get("variant") == "brute" and kw.get("num_jk", 0) > 0:
			assert box.last_stats["kernel"] == 2, "variance calls use the ordered kernels"
			assert any(k.endswith("_sigmasq") and np.any(v != 0) for k, v in want.items())
	if "nan_rule" in name:
		# the reference's NaN rule (measure_w_box_jk.py:411-417, measure_m_box_jk.py:431-438) must actually fire, in both
		# kernels, on exactly the pairs the oracle zeroes
		import pyoracle
		pos, pos_s, axis, e, w, w_s = pyoracle.prepare({**data, "weight": np.ones(len(data["Position"])),
														"weight_shape_sample": np.ones(len(data["Position"]))})
		r_bins, pi_bins, mu_bins = pyoracle.make_bins((0.1, 20.0), kw["num_bins_r"], kw["num_bins_pi"], None,
													   meta["catalogue"]["boxsize"])
		geom = "rppi" if kw["kind"] == "w" else "rmu"
		ref = pyoracle.paircount(geom, pos, w, None, pos_s, axis, e, w_s, None, r_bins, (0.1, 20.0),
								 pi_bins if geom == "rppi" else mu_bins, meta["catalogue"]["boxsize"], True, int(data["LOS"]), 1.0)
		assert ref["n_nan"] >= 50
		assert box.last_stats["nan_rule"] == ref["n_nan"], (box.last_stats["nan_rule"], ref["n_nan"])
	dd_key = [k for k in want if k.endswith("xi_gg/All_DD")][0]
	if not meta["catalogue"].get("weights"):
		assert box.last_stats["binned"] == int(want[dd_key].sum())
		assert np.array_equal(box.last_result["count"], want[dd_key].astype(np.int64))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("geom,kind", [("rppi", "w"), ("rmu", "multipoles")])
def test_against_oracle_100k(torch_cuda, oracle, tmp_path, geom, kind, kernel):
	"""N = 1e5 (3e8 / 4e7 pairs): far beyond what the Python reference can do; oracle = C restatement."""
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(100000, 205.0, seed=77, weights=(geom == "rmu"))
	out = str(tmp_path / "o.hdf5")
	box = MeasureIABox(data, out, boxsize=205.0, num_bins_r=10, num_bins_pi=8)
	box.kernel = kernel
	(box.measure_xi_w if kind == "w" else box.measure_xi_multipoles)("All", "both", num_jk=27, temp_file_path=False)
	want = oracle.measure(data, kind, num_jk=27, boxsize=205.0, num_bins_r=10, num_bins_pi=8,
						  n_threads=oracle.max_threads(), variant="brute")  # temp_file_path=False: `_sigmasq` is accumulated
	count = want.pop("__meta__/count")
	want.pop("__meta__/n_tested")
	assert np.array_equal(box.last_result["count"], count)
	assert np.any(want[("w" if kind == "w" else "multipoles") + "/xi_g_plus/All_sigmasq"] > 0)
	pu.assert_datasets_match(read_all(out), want, exact_counts=(geom == "rppi"), label=f"{geom}[{kernel}]: ")
	if kernel != "general":
		assert box.last_stats["kernel"] in TILED


def _host_call(torch, oracle, data, geom, num_jk, boxsize, n_r, n_2, kernel=0, shard=(0, 1)):
	"""mia_paircount_host through ctypes only (host numpy buffers in / out)."""
	from measure_ia_b200 import MeasureIABox, ops
	box = MeasureIABox(data, None, boxsize=boxsize, num_bins_r=n_r, num_bins_pi=n_2)
	pos, pos_s, axis, e, w, w_s, same = box._prepare(None, "distortion")
	if pos is pos_s and np.array_equal(w, w_s):
		w_s = w  # the C ABI recognises an auto-correlation by pointer identity of pos / weight / jk (include/mia_b200.h)
	L = round(num_jk ** (1 / 3)) if num_jk else 0
	jk = box._jackknife_labels(pos, L).astype(np.int32) if num_jk else None
	r2_thr, thr2, rp2_cut, _ = box._thresholds_for(geom, None)
	P = ops.MiaParams(ops.MIA_ABI_VERSION, 0 if geom == "rppi" else 1, n_r, n_2, int(data["LOS"]), 1, num_jk, kernel, boxsize,
					  float(box.r_bins[-1]), rp2_cut, r2_thr.ctypes.data, thr2.ctypes.data, None)
	ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
	D = ops.MiaSample(len(pos), ptr(pos), ptr(w), ptr(jk), None, None)
	S = ops.MiaSample(len(pos_s), ptr(pos_s), ptr(w_s), ptr(jk), ptr(axis), ptr(e))
	nb = (n_r, n_2)
	o = dict(dd_count=np.zeros(nb, np.int64), dd_w=np.zeros(nb), spd=np.zeros(nb), scd=np.zeros(nb),
			 dd_jk_count=np.zeros((num_jk,) + nb, np.int64), dd_jk_w=np.zeros((num_jk,) + nb),
			 spd_jk=np.zeros((num_jk,) + nb), stats=np.zeros(8, np.uint64))
	H = ops.MiaHist(*[ptr(o[k]) if o[k].size else None for k in ("dd_count", "dd_w", "spd", "scd", "dd_jk_count",
																   "dd_jk_w", "spd_jk", "stats")])
	ops.paircount_host(P, D, S, H, shard=shard, device=0)
	R, _ = box._responsivity(w_s, e)
	ref = oracle.paircount(geom, pos, w, jk, pos_s, axis, e, w_s, jk, box.r_bins, (0.1, 20.0),
						   box.pi_bins if geom == "rppi" else box.mu_r_bins, boxsize, True, int(data["LOS"]), 1.0,
						   num_box=num_jk, n_threads=oracle.max_threads())
	return o, ref


@pytest.mark.parametrize("geom", ["rppi", "rmu"])
def test_c_abi_host_entry(torch_cuda, oracle, geom):
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(20000, 150.0, seed=5, weights=True, los=1)
	o, ref = _host_call(torch_cuda, oracle, data, geom, 8, 150.0, 6, 12)
	assert np.array_equal(o["dd_count"], ref["count"])
	assert int(o["stats"][1]) == int(ref["count"].sum())
	for a, b in ((o["dd_w"], ref["DD"]), (o["spd"], ref["SpD"]), (o["scd"], ref["ScD"]), (o["dd_jk_w"], ref["DD_jk"]),
				 (o["spd_jk"], ref["SpD_jk"])):
		np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-11 * np.abs(b).max())


@pytest.mark.parametrize("geom,plan", [("rppi", "default"), ("rppi", "rows"), ("rmu", "default")])
def test_shards_sum_to_whole(torch_cuda, oracle, monkeypatch, geom, plan):
	"""Multi-GPU partitioning property on one GPU: the shards of the shape sample add up to the unsharded result (default
	plan: cell-by-cell kernel, static slots; rows: the symmetric kernel, interleaved slots handed out dynamically)."""
	from measure_ia_b200.synthetic import uniform_box
	if plan == "rows":
		monkeypatch.setenv("MIA_RPPI_V2", "2")
	data = uniform_box(30000, 205.0, seed=9)
	whole, _ = _host_call(torch_cuda, oracle, data, geom, 27, 205.0, 10, 8)
	if plan == "rows" or geom == "rmu":
		assert int(whole["stats"][4]) == 4  # the symmetric kernels (same arrays on both sides through the host entry point)
	parts = [_host_call(torch_cuda, oracle, data, geom, 27, 205.0, 10, 8, shard=(i, 3))[0] for i in range(3)]
	assert np.array_equal(sum(p["dd_count"] for p in parts), whole["dd_count"])
	assert np.array_equal(sum(p["dd_jk_count"] for p in parts), whole["dd_jk_count"])
	np.testing.assert_allclose(sum(p["spd"] for p in parts), whole["spd"], rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("temp", [False, "tmp/"], ids=["brute_semantics", "tree_semantics"])
def test_exact_quarter_scaling_with_half_weights(torch_cuda, tmp_path, monkeypatch, temp):
	"""Reference tests/test_weights.py:34-35: weights 0.5 scale DD and w_g+ by EXACTLY 1/4 (needs run-to-run
	deterministic accumulation order), covariances by 1/16."""
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(20000, 205.0, seed=31)
	monkeypatch.setenv("MIA_RPPI_V2", "2")  # row-streaming kernels also for this sparse catalogue
	out = str(tmp_path / "w.hdf5")
	box = MeasureIABox(data, out, boxsize=205.0, num_bins_r=10, num_bins_pi=8)
	box.measure_xi_w("A", "both", 8, temp_file_path=temp)
	assert box.last_stats["kernel"] == (2 if temp is False else 4)  # ordered kernel with variance / symmetric kernel
	box.data["weight"] = np.array([0.5] * 20000)
	box.data["weight_shape_sample"] = np.array([0.5] * 20000)
	box.measure_xi_w("B", "both", 8, temp_file_path=temp)
	assert box.last_stats["kernel"] == (2 if temp is False else 4)
	g = read_all(out)
	assert np.array_equal(g["w/xi_gg/A_DD"], 4 * g["w/xi_gg/B_DD"])
	if box.last_stats["kernel"] in TILED:  # the tiled kernels accumulate in a fixed order
		assert np.array_equal(g["w_g_plus/A"], 4 * g["w_g_plus/B"])
	np.testing.assert_allclose(g["w_g_plus/A"], 4 * g["w_g_plus/B"], rtol=1e-12)
	np.testing.assert_allclose(g["w_g_plus/A_jackknife_cov_8"], 16 * g["w_g_plus/B_jackknife_cov_8"], rtol=1e-9)
	np.testing.assert_allclose(g["w_gg/A_jackknife_cov_8"], 16 * g["w_gg/B_jackknife_cov_8"], rtol=1e-9)


def test_corr_type_invariance_and_symmetry(torch_cuda, tmp_path):
	"""Reference tests/test_w_sim_internal_consistency.py:35-54: 'g+', 'gg' and 'both' give identical numbers; plus the
	auto-correlation symmetry DD(r_p, Pi) == DD(r_p, -Pi)."""
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(30000, 205.0, seed=41)
	out = str(tmp_path / "c.hdf5")
	box = MeasureIABox(data, out, boxsize=205.0, num_bins_r=10, num_bins_pi=8)
	box.measure_xi_w("both", "both", 0, temp_file_path=False)
	box.measure_xi_w("gp", "g+", 0, temp_file_path=False)
	box.measure_xi_w("gg", "gg", 0, temp_file_path=False)
	g = read_all(out)
	assert np.array_equal(g["w_gg/both"], g["w_gg/gg"]) or np.allclose(g["w_gg/both"], g["w_gg/gg"], rtol=1e-13)
	np.testing.assert_allclose(g["w_g_plus/both"], g["w_g_plus/gp"], rtol=1e-10, atol=1e-14)
	assert "w_gg/gp" not in g and "w_g_plus/gg" not in g
	dd = g["w/xi_gg/both_DD"]
	assert np.array_equal(dd, dd[:, ::-1])


def test_out_of_box_coordinates_fail_loudly(torch_cuda):
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(1000, 100.0, seed=3)
	data["Position"] = data["Position"].copy()
	data["Position"][7, 1] = 100.0  # == boxsize: scipy's periodic KDTree raises in the reference
	box = MeasureIABox(data, None, boxsize=100.0)
	with pytest.raises(RuntimeError, match="outside"):
		box.measure_xi_w("x", "both", 0, temp_file_path=False)


def test_empty_and_tiny_inputs(torch_cuda, oracle, tmp_path):
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	for n in (1, 2, 33):
		data = uniform_box(n, 50.0, seed=n)
		box = MeasureIABox(data, str(tmp_path / f"t{n}.hdf5"), boxsize=50.0, num_bins_r=4, num_bins_pi=4)
		for kind, run in (("w", box.measure_xi_w), ("multipoles", box.measure_xi_multipoles)):
			run("All", "both", 8, temp_file_path=False)
			want = oracle.measure(data, kind, num_jk=8, boxsize=50.0, num_bins_r=4, num_bins_pi=4)
			assert np.array_equal(box.last_result["count"], want["__meta__/count"])


def test_device_preparation_matches_host(torch_cuda):
	"""The torch/GPU input preparation (box._prepare_device) against the numpy restatement of the reference's host code
	(box._prepare, _jackknife_labels, _responsivity): labels, axis and e bit-identical, responsivities to rounding."""
	import torch
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(50000, 90.0, seed=17, n_shape=30000, weights=True)
	data["Position"][:40, 0] = 30.0  # points exactly on jackknife faces keep label 0
	data["Position"][40:60, 2] = 0.0
	data["Position_shape_sample"][:25, 1] = 60.0
	rng = np.random.default_rng(3)
	masks = {"Position": rng.random(50000) < 0.8, "Position_shape_sample": rng.random(30000) < 0.7}
	masks["Axis_Direction"] = masks["q"] = masks["Position_shape_sample"]
	for use_masks in (None, masks):
		box = MeasureIABox(data, None, boxsize=90.0)
		m_host = None if use_masks is None else dict(use_masks)
		m_dev = None if use_masks is None else dict(use_masks)
		pos, pos_s, axis, e, w, w_s, same = box._prepare(m_host, "distortion")
		jk_p, jk_s = box._jackknife_labels(pos, 3), box._jackknife_labels(pos_s, 3)
		R, R_jk = box._responsivity(w_s, e, jk_s, 27)
		P = box._prepare_device(m_dev, "distortion", 3, torch.device("cuda", 0))
		assert np.array_equal(P["pos"].cpu().numpy(), pos) and np.array_equal(P["pos_s"].cpu().numpy(), pos_s)
		assert np.array_equal(P["axis"].cpu().numpy(), axis) and np.array_equal(P["e"].cpu().numpy(), e)
		assert np.array_equal(P["w"].cpu().numpy(), w) and np.array_equal(P["w_s"].cpu().numpy(), w_s)
		assert np.array_equal(P["jk_p"].cpu().numpy(), jk_p) and np.array_equal(P["jk_s"].cpu().numpy(), jk_s)
		assert (jk_p[:40] == 0).all() or use_masks is not None
		np.testing.assert_allclose(P["R"], R, rtol=1e-13)
		np.testing.assert_allclose(P["R_jk"], R_jk, rtol=1e-13)
		assert np.array_equal(P["n_p_k"], len(pos) - np.bincount(jk_p, minlength=27))
		assert np.array_equal(P["n_s_k"], len(pos_s) - np.bincount(jk_s, minlength=27))


@pytest.mark.parametrize("kernel", KERNELS)
def test_clustered_weighted_default_bins_vs_oracle(torch_cuda, oracle, tmp_path, kernel):
	"""Constructor-default 8 x 20 bins, half of the galaxies in Gaussian blobs (load balance, dense cells that split into
	many chunks), weights, cross-correlation of two different samples, 64 jackknife regions."""
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(120000, 300.0, seed=55, n_shape=60000, weights=True, clustered=0.5)
	out = str(tmp_path / "c.hdf5")
	box = MeasureIABox(data, out, boxsize=300.0)
	box.kernel = kernel
	box.measure_xi_w("All", "both", num_jk=64, temp_file_path=False)
	want = oracle.measure(data, "w", num_jk=64, boxsize=300.0, n_threads=oracle.max_threads(), variant="brute")
	count = want.pop("__meta__/count")
	want.pop("__meta__/n_tested")
	assert np.array_equal(box.last_result["count"], count)
	pu.assert_datasets_match(read_all(out), want, exact_counts=False, label=f"clustered[{kernel}]: ")
	if kernel != "general":
		assert box.last_stats["kernel"] in TILED


@pytest.mark.parametrize("temp", [False, "tmp/"], ids=["ordered_with_variance", "symmetric"])
def test_tiled_kernel_is_bit_reproducible(torch_cuda, monkeypatch, temp):
	"""Two runs of the tiled kernel give identical bits for every fp64 sum (fixed accumulation order)."""
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	data = uniform_box(60000, 205.0, seed=61, weights=True)
	monkeypatch.setenv("MIA_RPPI_V2", "2")
	box = MeasureIABox(data, None, boxsize=205.0, num_bins_r=10, num_bins_pi=8)
	box.kernel = "tiled"
	box.measure_xi_w("a", "both", 27, temp_file_path=temp)
	assert box.last_stats["kernel"] == (2 if temp is False else 4)
	first = {k: np.array(v) for k, v in box.last_result.items() if isinstance(v, np.ndarray)}
	box.measure_xi_w("a", "both", 27, temp_file_path=temp)
	for k, v in first.items():
		assert np.array_equal(v, box.last_result[k]), k


# ---- tiled kernels against the general kernel over awkward configurations -------------------------------------------------
_CROSS_CASES = [
	# n, boxsize, seed, num_jk, n_r, n_2, generator / constructor options
	(3000, 205.0, 1, 27, 10, 8, {}),
	(20000, 150.0, 4, 8, 6, 12, dict(los=1, weights=True)),
	(5000, 50.0, 5, 8, 4, 4, {}),                 # tiny box: every column is a neighbour, chunks straddle +-L/2
	(2000, 30.0, 6, 27, 5, 5, {}),                # r_max > L/2
	(33, 50.0, 8, 8, 4, 4, {}),
	(60000, 300.0, 10, 64, 8, 20, dict(n_shape=30000, weights=True, clustered=0.5)),
	(30000, 205.0, 14, 27, 10, 8, dict(periodicity=False)),
	(30000, 205.0, 13, 125, 10, 10, dict(los=0)),
	(50000, 205.0, 15, 27, 10, 8, dict(pi_max=30.0)),
	# exact edges (lattice coordinates: Pi on bin edges and on +-L/2, dz = 0, r_p = r_min) and the NaN rule
	(16, 40.0, 14, 8, 5, 8, dict(gen="lattice", n_random=300, separation_limits=(2.5, 15.0))),
	(20, 50.0, 16, 27, 6, 10, dict(gen="lattice", los=0, separation_limits=(2.5, 20.0))),
	(12, 30.0, 15, 27, 4, 6, dict(gen="lattice", los=1, n_random=150, separation_limits=(2.5, 12.5), pi_max=7.5)),
	(20000, 100.0, 12, 8, 10, 8, dict(gen="aligned_pairs")),
	# 11^3 regions on a grid that cannot be aligned with them: rows whose region holds several labels take the cell-by-cell
	# path of the (symmetric) rows kernels, and the accumulator workspace cap (fewer worker slots) applies
	(5000, 60.0, 21, 1331, 6, 6, {}),
]


def _run_cross(kind, case, kernel):
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import GENERATORS
	n, L, seed, jk, n_r, n_2, kw = case
	kw = dict(kw)
	per = kw.pop("periodicity", True)
	pi_max = kw.pop("pi_max", None)
	limits = list(kw.pop("separation_limits", (0.1, 20.0)))
	data = GENERATORS[kw.pop("gen", "uniform")](n, L, seed=seed, **kw)
	box = MeasureIABox(data, None, boxsize=L, separation_limits=limits, num_bins_r=n_r, num_bins_pi=n_2, periodicity=per,
					   pi_max=pi_max)
	box.kernel = kernel
	# a temporary path = the reference's tree variants (no `_sigmasq` accumulation): auto-correlations may then take the
	# symmetric kernel; temp_file_path=False (brute variants, variance accumulated) is covered by the fixture / oracle tests
	(box.measure_xi_w if kind == "w" else box.measure_xi_multipoles)("a", "both", jk, temp_file_path="tmp/")
	return box.last_result, box.last_stats


def _assert_same_sums(got, want, label, noise=None):
	assert np.array_equal(got["count"], want["count"]), f"{label}: pair counts differ"
	assert np.array_equal(got["count_jk"], want["count_jk"]), f"{label}: jackknife pair counts differ"
	# absolute floor: a bin sums `count` terms of magnitude <= ~1 in different orders, so sums that cancel exactly in exact
	# arithmetic (S+D / SxD on a perfect lattice) carry ~eps * count of rounding noise and nothing else; the general kernel's
	# unordered atomic accumulation measured ~1e-15 x (pairs in the bin) against the fixed-order kernels.  `noise`: measured
	# run-to-run differences of the general kernel, per array
	floor = 4e-15 * float(np.asarray(want["count"]).max(initial=0))
	for k in ("DD", "SpD_raw", "ScD_raw", "DD_jk", "SpD_jk"):
		a, b = np.asarray(want[k]), np.asarray(got[k])
		if a.size:
			tol = pu.RTOL * np.abs(a) + pu.ATOL_SCALE * np.abs(a).max() + floor + (2.0 * noise[k] if noise else 0.0)
			assert (np.abs(a - b) <= tol).all(), f"{label}: {k} differs by {np.abs(a - b).max():.3e}"


@pytest.mark.parametrize("mode", ["default", "rows", "rows_ordered", "cells", "split"])
@pytest.mark.parametrize("case", _CROSS_CASES, ids=[f"n{c[0]}_L{int(c[1])}_jk{c[3]}" for c in _CROSS_CASES])
@pytest.mark.parametrize("kind", ["w", "multipoles"])
def test_tiled_kernels_match_general(torch_cuda, monkeypatch, kind, case, mode):
	"""Every tiled code path (row-streaming / cell-by-cell (r_p, Pi), column-streaming (r, mu_r), tasks cut into parts as on
	many GPUs) against the reference-exact general kernel: pair counts bit-identical, sums to 1e-10."""
	if kind == "multipoles" and (mode in ("rows", "rows_ordered", "cells") or case[6].get("pi_max")):
		pytest.skip("(r_p, Pi)-only variation")
	want, st_g = _run_cross(kind, case, "general")
	assert st_g["kernel"] == 1
	if mode in ("rows", "rows_ordered"):
		monkeypatch.setenv("MIA_RPPI_V2", "2")
	elif mode == "cells":
		monkeypatch.setenv("MIA_RPPI_V2", "0")
	elif mode == "split":
		monkeypatch.setenv("MIA_TASKS_PER_WARP", "1000")
	got, st_t = _run_cross(kind, case, "tiled_ordered" if mode == "rows_ordered" else "tiled")
	assert st_t["kernel"] in TILED
	if mode == "rows_ordered":
		assert st_t["kernel"] == 2
	if mode == "rows" and kind == "w" and "n_shape" not in case[6] and case[0] in (3000, 20000, 30000, 50000):
		# auto-correlations on grids wide enough for the half-space rule take the symmetric kernel (every unordered pair
		# visited once, both orderings accumulated: mia_tiled_rppi2s.cuh)
		assert st_t["kernel"] == 4, case
	if mode == "default" and kind == "multipoles" and "n_shape" not in case[6] and case[0] in (3000, 30000) and case[5] == 8:
		assert st_t["kernel"] == 4, case  # (r, mu_r): symmetric variant of the column-streaming kernel
	assert st_t["binned"] == int(want["count"].sum())
	assert st_t["nan_rule"] == st_g["nan_rule"]
	if case[6].get("gen") == "aligned_pairs":
		assert st_g["nan_rule"] >= 500
	_assert_same_sums(got, want, f"{kind}/{mode}")


@pytest.mark.parametrize("workload", ["cfg2", "cfg3"])
def test_full_size_tiled_matches_general(torch_cuda, workload):
	"""BASELINE.json configs[1] / [2] at FULL size (1e6 galaxies, L = 205, 10 x 8 bins, 27 regions): the plan the bench
	runs (rows at r_max / 10 for (r_p, Pi); the 3-D column grid for (r, mu_r)) against the reference-exact general kernel.
	Pair counts and jackknife pair counts bit-identical, sums to 1e-10; the NaN-rule pairs (|c| > 1 by rounding,
	measure_w_box_jk.py:411-417) agree and do occur at this size."""
	case = (1_000_000, 205.0, 1, 27, 10, 8, {})
	kind = "w" if workload == "cfg2" else "multipoles"
	want, st_g = _run_cross(kind, case, "general")
	want2, _ = _run_cross(kind, case, "general")  # the checker's own noise: unordered atomic adds of 3e10 / 4e9 terms
	noise = {k: float(np.abs(np.asarray(want[k]) - np.asarray(want2[k])).max()) for k in ("DD", "SpD_raw", "ScD_raw", "DD_jk", "SpD_jk")}
	got, st_t = _run_cross(kind, case, "auto")
	assert st_g["kernel"] == 1 and st_t["kernel"] == 4  # both geometries: the symmetric auto-correlation kernels
	assert st_t["binned"] == st_g["binned"] == int(want["count"].sum())
	assert st_t["nan_rule"] == st_g["nan_rule"]
	print(f"{workload}: {st_t['binned']} pairs, nan_rule {st_t['nan_rule']}, tested {st_t['tested']}, general-kernel noise {noise}")
	_assert_same_sums(got, want, f"{workload} full size", noise=noise)
	# the ordered tiled kernel: an independent pair loop (every ordered pair on its own), also with fixed-order sums
	ordered, st_o = _run_cross(kind, case, "tiled_ordered")
	assert st_o["kernel"] == 2 and st_o["tested"] > 1.5 * st_t["tested"]
	assert np.array_equal(ordered["count"], got["count"]) and np.array_equal(ordered["count_jk"], got["count_jk"])
	for k in ("SpD_raw", "ScD_raw", "SpD_jk"):
		a, b = np.asarray(ordered[k]), np.asarray(got[k])
		assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max(), f"{workload}: symmetric vs ordered kernel, {k}: {np.abs(a - b).max():.3e}"
	dd = got["count"]
	if kind == "w":
		assert np.array_equal(dd, dd[:, ::-1]), "auto-correlation: DD(r_p, Pi) == DD(r_p, -Pi)"

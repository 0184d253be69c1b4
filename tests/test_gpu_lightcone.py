"""GPU: the light-cone brute pair-loop operator (mia_lightcone.cuh) through the product's public API and the
C ABI, against (1) every dataset of the reference's output files (tests/golden/lc_*.npz), (2) the numpy oracle
(oracle/pylightcone.py) on a larger seeded catalogue incl. the per-patch jackknife sums, (3) itself: shards add up, host entry =
device entry, unsorted input fails loudly.  Bar: pair counts bit-exact, every sum within 1e-10 (parity_util)."""
import ctypes

import numpy as np
import pytest

import parity_util as pu
from test_lightcone_host import LC_FIXTURES, compare_with_fixture, load_lc_fixture, run_product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
	import torch
	if not torch.cuda.is_available():
		pytest.skip("GPU tests need a CUDA device")
	return torch


def _read_all(path):
	from test_host_post import read_all
	return read_all(path)


@pytest.mark.parametrize("name", LC_FIXTURES)
def test_reference_lightcone_fixture(torch_cuda, tmp_path, name):
	meta, want = load_lc_fixture(name)
	out = str(tmp_path / "lc.hdf5")
	obj = run_product(meta, out)
	assert obj.last_stats["kernel"] == 5 and obj.last_stats["launches"] >= 3 and obj.last_stats["thresholds_clean"]
	compare_with_fixture(_read_all(out), want, f"{name}: ", not meta["catalogue"].get("weights"))


def test_device_distances_equal_numpy_bit_for_bit(torch_cuda):
	"""The distance integral runs on the device (lightcone._pair_sums): +, *, /, sqrt in the same order as the numpy version the
	reference fixtures were generated with -- the bits must agree, or pairs next to a bin edge could move."""
	from measure_ia_b200 import cosmo
	c = cosmo.Cosmology()
	om = c["Omega_c"] + c["Omega_b"]
	for zmax in (0.4, 3.0, 12.0):  # 16 / 32 / 64 nodes
		z = np.random.default_rng(7).uniform(0.0, zmax, 200_000)
		a = 1 / (1 + z)
		dev = cosmo.flat_lcdm_distance(om, c["h"], 1 / (1 + torch_cuda.from_numpy(z).cuda())).cpu().numpy()
		assert np.array_equal(dev, cosmo.comoving_radial_distance(c, a)), zmax


def _catalogue(n, ns, seed, K=6):
	rng = np.random.default_rng(seed)
	d = {"RA": rng.uniform(10.0, 16.0, n), "DEC": rng.uniform(-3.0, 3.0, n), "Redshift": rng.uniform(0.10, 0.16, n),
		 "RA_shape_sample": rng.uniform(10.0, 16.0, ns), "DEC_shape_sample": rng.uniform(-3.0, 3.0, ns),
		 "Redshift_shape_sample": rng.uniform(0.10, 0.16, ns), "e1": rng.normal(0, 0.2, ns), "e2": rng.normal(0, 0.2, ns),
		 "weight": rng.uniform(0.5, 1.5, n), "weight_shape_sample": rng.uniform(0.5, 1.5, ns)}
	pp = np.minimum((d["RA"] - 10.0) / 6.0 * K, K - 1).astype(int) + 1  # patch labels 1..K, as kmeans-style labels may start at 1
	ps = np.minimum((d["RA_shape_sample"] - 10.0) / 6.0 * K, K - 1).astype(int) + 1
	return d, (pp, ps)


def _oracle_sums(obj, geom, shapes, over_h, rp_cut, patches):
	from test_lightcone_host import oracle_lc_pair_sums
	return oracle_lc_pair_sums()(obj, geom, shapes, None, over_h, None, rp_cut, patches)


@pytest.mark.parametrize("geom,shapes,over_h,rp_cut", [("rppi", True, False, None), ("rppi", False, True, None),
														("rmu", True, True, 1.0), ("rmu", False, False, None)])
def test_operator_matches_oracle_with_patches(torch_cuda, geom, shapes, over_h, rp_cut):
	from measure_ia_b200.lightcone import MeasureIALightcone
	data, patches = _catalogue(7000, 5000, seed=31)
	obj = MeasureIALightcone(data, None, None, [0.3, 25.0], 9, 8, 60.0)
	got = obj._pair_sums(geom, shapes, None, over_h, None, rp_cut, patches)
	stats = obj.last_stats
	want = _oracle_sums(obj, geom, shapes, over_h, rp_cut, patches)
	assert stats["binned"] == int(want["count"].sum()) > 100_000 and stats["tested"] <= 7000 * 5000
	assert np.array_equal(got["count"], want["count"]), "pair counts differ from the oracle"
	assert np.array_equal(got["touch_count"], want["touch_count"]), "per-patch pair counts differ from the oracle"
	assert got["touch_count"].shape == (6, 9, 8)
	for k in ("DD", "SpD", "ScD", "touch_DD", "touch_SpD"):
		a, b = want[k], got[k]
		scale = np.abs(a).max()
		assert np.all(np.abs(a - b) <= 1e-10 * np.abs(a) + 1e-11 * scale), (k, float(np.abs(a - b).max()), scale)
	if not shapes:
		assert not got["SpD"].any() and not got["ScD"].any()
	if geom == "rppi":  # the chi window really culls: far fewer separations than the full N_p x N_s
		assert stats["tested"] < 0.8 * 7000 * 5000


def test_shards_host_entry_and_unsorted_input(torch_cuda):
	torch = torch_cuda
	from measure_ia_b200 import ops
	from measure_ia_b200.lightcone import MeasureIALightcone
	data, patches = _catalogue(3000, 2500, seed=32, K=3)
	obj = MeasureIALightcone(data, None, None, [0.3, 25.0], 7, 6, 50.0)
	whole = obj._pair_sums("rppi", True, None, False, None, None, patches)
	# ---- the same call, position sample in two shards, straight through the operator ------------------------------------------
	from measure_ia_b200 import cosmo
	c = cosmo.Cosmology()
	f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
	chi_p = cosmo.comoving_radial_distance(c, 1 / (1 + data["Redshift"]))
	chi_s = cosmo.comoving_radial_distance(c, 1 / (1 + data["Redshift_shape_sample"]))
	e1, e2 = data["e1"], data["e2"]
	phi = np.arctan2(np.sin(0.5 * np.arctan2(e2, e1)), np.cos(0.5 * np.arctan2(e2, e1)))
	e = np.sqrt(e1 ** 2 + e2 ** 2)
	pos = dict(ra=f(data["RA"]), dec=f(data["DEC"]), chi=f(chi_p), cosdec=np.cos(f(data["DEC"]) / 180 * np.pi), weight=f(data["weight"]),
			   patch=(patches[0] - 1).astype(np.int32))
	shp = dict(ra=f(data["RA_shape_sample"]), dec=f(data["DEC_shape_sample"]), chi=f(chi_s), weight=f(data["weight_shape_sample"]),
			   e1=e * np.cos(2 * phi), e2=e * np.sin(2 * phi), patch=(patches[1] - 1).astype(np.int32))
	op, os_ = np.argsort(pos["chi"], kind="stable"), np.argsort(shp["chi"], kind="stable")
	pos = {k: np.ascontiguousarray(v[op]) for k, v in pos.items()}
	shp = {k: np.ascontiguousarray(v[os_]) for k, v in shp.items()}
	up = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}  # noqa: E731
	r2_thr, thr2, _, _ = obj._thresholds_for("rppi", None)
	args = (torch.from_numpy(r2_thr), torch.from_numpy(thr2), ops.GEOM_RPPI, True, 3, 1.0, 0.0)
	parts = [ops.lightcone_paircount(up(pos), up(shp), *args, i, 2) for i in range(2)]
	assert np.array_equal((parts[0][0] + parts[1][0]).cpu().numpy(), whole["count"])
	assert np.array_equal((parts[0][4] + parts[1][4]).cpu().numpy(), whole["touch_count"])
	assert parts[0][0].sum() > 0 and parts[1][0].sum() > 0
	for i, k in ((1, "DD"), (2, "SpD"), (3, "ScD"), (5, "touch_DD"), (6, "touch_SpD")):
		a = (parts[0][i] + parts[1][i]).cpu().numpy()
		assert np.allclose(a, whole[k], rtol=1e-10, atol=1e-11 * np.abs(whole[k]).max()), k
	# ---- host entry of the C ABI (what a binding without a device-array library calls) -------------------------------------------
	n_r, n_2 = 7, 6
	ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
	params = ops.MiaLcParams(ops.MIA_ABI_VERSION, ops.GEOM_RPPI, n_r, n_2, 3, 1, 1.0, 0.0, ptr(r2_thr), ptr(thr2), None)
	D = ops.MiaLcSample(len(pos["ra"]), ptr(pos["ra"]), ptr(pos["dec"]), ptr(pos["chi"]), ptr(pos["cosdec"]), ptr(pos["weight"]), None, None,
						ptr(pos["patch"]))
	S = ops.MiaLcSample(len(shp["ra"]), ptr(shp["ra"]), ptr(shp["dec"]), ptr(shp["chi"]), None, ptr(shp["weight"]), ptr(shp["e1"]),
						ptr(shp["e2"]), ptr(shp["patch"]))
	cnt, jc = np.zeros((n_r, n_2), dtype=np.int64), np.zeros((3, n_r, n_2), dtype=np.int64)
	ddw, spd, scd = (np.zeros((n_r, n_2)) for _ in range(3))
	jw, jsp = (np.zeros((3, n_r, n_2)) for _ in range(2))
	stats = np.zeros(8, dtype=np.uint64)
	H = ops.MiaHist(ptr(cnt), ptr(ddw), ptr(spd), ptr(scd), ptr(jc), ptr(jw), ptr(jsp), ptr(stats), None)
	ops.lightcone_paircount_host(params, D, S, H, device=torch.cuda.current_device())
	assert np.array_equal(cnt, whole["count"]) and np.array_equal(jc, whole["touch_count"]) and int(stats[4]) == 5
	assert np.allclose(spd, whole["SpD"], rtol=1e-10, atol=1e-11 * np.abs(whole["SpD"]).max())
	# ---- a sample that is not sorted by chi is refused (the window cull would silently drop pairs) ---------------------------------
	bad = dict(pos)
	bad["chi"] = np.ascontiguousarray(pos["chi"][::-1])
	with pytest.raises(RuntimeError, match="sorted"):
		ops.lightcone_paircount(up(bad), up(shp), *args)
	# ---- empty samples -----------------------------------------------------------------------------------------------------------
	empty = {k: v[:0] for k, v in pos.items()}
	out = ops.lightcone_paircount(up(empty), up(shp), *args)
	assert int(out[0].sum()) == 0 and int(out[7][4]) == 5

"""CPU: the light-cone path without a GPU.

(1) the numpy oracle of the four brute pair loops (oracle/pylightcone.py) together with the product's host side
    (measure_ia_b200/lightcone.py: input selection, the reference's temporary data dictionaries S+D / S+R / SD / SR / RD / RR,
    estimators for clusters and galaxies, w and multipoles, HDF5 layout) against every dataset the UNMODIFIED reference wrote
    (tests/golden/lc_*.npz, oracle/make_golden_lightcone.py) -- the pair sums are injected from the oracle (test-only);
(2) the oracle's jackknife "touch" sums against re-running it without each patch (the reference's leave-one-patch-out rule,
    measure_jackknife.py:116-134);
(3) error behaviour and the C-ABI symbols of the light-cone entry points."""
import json
import os

import numpy as np
import pytest

import parity_util as pu
from measure_ia_b200 import h5lite
from measure_ia_b200.lightcone import MeasureIALightcone
from test_host_post import read_all

LC_FIXTURES = sorted(f[:-4] for f in os.listdir(pu.GOLDEN) if f.startswith("lc_") and f.endswith(".npz"))


def load_lc_fixture(name):
	z = np.load(os.path.join(pu.GOLDEN, name + ".npz"))
	meta = json.loads(str(z["__meta__"]))
	return meta, {k.replace("|", "/"): z[k] for k in z.files if k != "__meta__"}


def oracle_lc_pair_sums():
	"""Stand-in for MeasureIALightcone._pair_sums: same selection, the sums from oracle/pylightcone.py."""
	import pylightcone

	def _pair_sums(self, geom, shapes, masks, over_h, cosmology, rp_cut=None, patches=None):
		sel = self._select(masks, shapes)
		pos, h = pylightcone.sample(sel["RA"], sel["DEC"], sel["Redshift"], sel["weight"], cosmology, over_h)
		shp, _ = pylightcone.sample(sel["RA_shape_sample"], sel["DEC_shape_sample"], sel["Redshift_shape_sample"],
									sel["weight_shape_sample"], cosmology, over_h, sel.get("e1"), sel.get("e2"))
		bins2 = self.pi_bins if geom == "rppi" else self.mu_r_bins
		kw, extra = {}, {}
		if patches is not None:
			pp, ps = np.asarray(patches[0]), np.asarray(patches[1])
			if masks is not None and len(pp) == len(masks["Redshift"]) and len(ps) == len(masks["Redshift_shape_sample"]):
				pp, ps = pp[masks["Redshift"]], ps[masks["Redshift_shape_sample"]]
			lo = int(min(np.min(pp), np.min(ps)))
			kw = dict(patches_pos=pp - lo, patches_shape=ps - lo, num_patches=int(max(np.max(pp), np.max(ps))) - lo + 1)
			extra = dict(patch_lo=lo, patches=(pp, ps))
		r = pylightcone.pair_sums(geom, pos, shp, self.r_min, self.r_max, self.r_bins, bins2, self.num_bins_r, self.num_bins_pi,
								  h=h, over_h=over_h, rp_cut=0.0 if rp_cut is None else rp_cut, shapes=shapes, **kw)
		r.update(extra)
		self.last_stats = dict(rank=0)
		self.last_result = r
		return r
	return _pair_sums


def run_product(meta, out, monkeypatch=None):
	import make_golden_lightcone as mg
	data, randoms, masks = mg.build_inputs(meta["catalogue"])
	jk = mg.build_patches(meta["catalogue"], data, randoms)
	b, call = meta["binning"], meta["call"]
	obj = MeasureIALightcone(data, randoms, out, b["separation_limits"], b["num_bins_r"], b["num_bins_pi"], b["pi_max"], 1)
	with np.errstate(all="ignore"):
		if call["kind"] == "w":
			obj.measure_xi_w(call["IA_estimator"], "All", call["corr_type"], jk_patches=jk, measure_cov=bool(jk), masks=masks,
							 over_h=call["over_h"])
		else:
			obj.measure_xi_multipoles(call["IA_estimator"], "All", call["corr_type"], jk_patches=jk, calc_errors=bool(jk), masks=masks,
									  over_h=call["over_h"], rp_cut=call.get("rp_cut"))
	return obj


def compare_with_fixture(got, want, label, unit_weights):
	assert set(got) == set(want), (sorted(set(got) ^ set(want)))
	# pair counts: exact for unit weights; every sum to 1e-10 (relative to the largest entry: random-sign terms cancel)
	pu.assert_datasets_match(got, want, exact_counts=False, label=label)
	if unit_weights:
		for k in want:
			if k.endswith(("_DD", "_RR", "_RD", "_SR")):
				assert np.array_equal(got[k], want[k]), f"{label}{k}: pair counts differ"


@pytest.mark.parametrize("name", LC_FIXTURES)
def test_host_pipeline_reproduces_reference_lightcone_files(tmp_path, monkeypatch, name):
	meta, want = load_lc_fixture(name)
	monkeypatch.setattr(MeasureIALightcone, "_pair_sums", oracle_lc_pair_sums())
	out = str(tmp_path / "lc.hdf5")
	run_product(meta, out)
	compare_with_fixture(read_all(out), want, f"{name}: ", not meta["catalogue"].get("weights"))


def test_oracle_touch_sums_are_the_leave_one_patch_out_rule():
	import pylightcone
	rng = np.random.default_rng(5)
	n, ns, K = 260, 190, 4
	ra, dec, z = rng.uniform(10, 13, n), rng.uniform(-1.5, 1.5, n), rng.uniform(0.10, 0.125, n)
	ra_s, dec_s, z_s = rng.uniform(10, 13, ns), rng.uniform(-1.5, 1.5, ns), rng.uniform(0.10, 0.125, ns)
	e1, e2 = rng.normal(0, 0.2, ns), rng.normal(0, 0.2, ns)
	pp, ps = np.minimum((ra - 10.0) / 3.0 * K, K - 1).astype(int), np.minimum((ra_s - 10.0) / 3.0 * K, K - 1).astype(int)
	r_bins = np.logspace(np.log10(0.5), np.log10(20.0), 6)
	for geom, bins2 in (("rppi", np.linspace(-40, 40, 7)), ("rmu", np.linspace(-1, 1, 7))):
		pos, h = pylightcone.sample(ra, dec, z, rng.uniform(0.5, 1.5, n))
		shp, _ = pylightcone.sample(ra_s, dec_s, z_s, rng.uniform(0.5, 1.5, ns), e1=e1, e2=e2)
		full = pylightcone.pair_sums(geom, pos, shp, 0.5, 20.0, r_bins, bins2, 5, 6, patches_pos=pp, patches_shape=ps, num_patches=K)
		assert full["count"].sum() > 1000
		for k in range(K):
			sub_p = {key: v[pp != k] for key, v in pos.items()}
			sub_s = {key: v[ps != k] for key, v in shp.items()}
			part = pylightcone.pair_sums(geom, sub_p, sub_s, 0.5, 20.0, r_bins, bins2, 5, 6)
			assert np.array_equal(full["count"] - full["touch_count"][k], part["count"]), (geom, k)
			for a, b in (("DD", "touch_DD"), ("SpD", "touch_SpD")):
				assert np.allclose(full[a] - full[b][k], part[a], rtol=1e-10, atol=1e-11 * np.abs(full[a]).max()), (geom, k, a)


def test_lightcone_error_behaviour(tmp_path):
	import make_golden_lightcone as mg
	data, randoms, _ = mg.build_inputs(dict(n=20, n_shape=20, n_rand=30, seed=1))
	obj = MeasureIALightcone(data, randoms, str(tmp_path / "x.hdf5"), [0.5, 20.0], 5, 6, 40.0)
	with pytest.raises(KeyError, match="IA_estimator"):
		obj.measure_xi_w("stars", "All", "both", measure_cov=False)
	with pytest.raises(KeyError, match="corr_type"):
		obj.measure_xi_w("galaxies", "All", "g++", measure_cov=False)
	with pytest.raises(ValueError, match="jk_patches or num_jk"):  # measure_cov / calc_errors default to True, as in the reference
		obj.measure_xi_multipoles("clusters", "All", "both")
	with pytest.raises(KeyError, match="randoms_jk"):  # the reference's estimator fails the same way for a 'gg'-only jackknife
		obj.measure_xi_w("galaxies", "All", "gg", jk_patches=mg.build_patches(dict(jk=3), data, randoms))
	with pytest.raises(ValueError, match="pi_max and boxsize"):
		MeasureIALightcone(data, randoms, None)
	import torch
	if not torch.cuda.is_available():  # no CPU fallback
		with pytest.raises(RuntimeError, match="CUDA"):
			obj.measure_xi_w("galaxies", "All", "both", measure_cov=False)


def test_flat_lcdm_distance_against_direct_quadrature():
	"""measure_ia_b200/cosmo.py (used when pyccl is absent): 64-point Gauss-Legendre vs scipy's adaptive quadrature."""
	from scipy.integrate import quad
	from measure_ia_b200 import cosmo
	c = cosmo.Cosmology()
	z = np.array([0.0, 0.05, 0.3, 1.0, 3.0])
	chi = cosmo.comoving_radial_distance(c, 1 / (1 + z))
	om = c["Omega_c"] + c["Omega_b"]
	want = [cosmo.C_KM_S / (100 * c["h"]) * quad(lambda x: 1 / np.sqrt(om * (1 + x) ** 3 + 1 - om), 0, zz, epsabs=0, epsrel=1e-13)[0]
			for zz in z]
	assert np.allclose(chi, want, rtol=1e-12, atol=0) and chi[0] == 0.0
	assert abs(chi[2] - 1202.3) < 0.5  # 1.20 Gpc at z = 0.3 for Om = 0.27, h = 0.7
	assert np.array_equal(cosmo.comoving_radial_distance(lambda a: 3000.0 * (1 / a - 1), 1 / (1 + z)), 3000.0 * (1 / (1 / (1 + z)) - 1))
	# the same function on torch tensors (the product evaluates it on the device, where sqrt and division are IEEE-exact and the
	# result equals numpy's bit for bit: tests/test_gpu_lightcone.py; torch's CPU sqrt is not correctly rounded, hence 4 ulp here)
	import torch
	zz = np.random.default_rng(1).uniform(0.01, 2.5, 5000)
	t = cosmo.flat_lcdm_distance(om, c["h"], torch.from_numpy(1 / (1 + zz))).numpy()
	assert np.allclose(t, cosmo.comoving_radial_distance(c, 1 / (1 + zz)), rtol=1e-15, atol=0)


def test_jackknife_patches_from_num_jk(tmp_path, monkeypatch):
	"""`num_jk` without `jk_patches` (measure_IA.py:415-418): patches from a spherical k-means of the position randoms
	(kmeans_radec's role, measure_IA_base.py:744-803) -- every sample labelled with its nearest centre, all labels used,
	reproducible; then the whole jackknife pipeline on those patches."""
	import make_golden_lightcone as mg
	data, randoms, _ = mg.build_inputs(dict(n=260, n_shape=220, n_rand=500, seed=33, weights=True))
	obj = MeasureIALightcone(data, randoms, str(tmp_path / "k.hdf5"), [0.5, 20.0], 5, 6, 40.0)
	obj._prepare_randoms()
	jk = obj.assign_jackknife_patches(data, randoms, 4)
	assert set(jk) == {"position", "shape", "randoms_position", "randoms_shape"}
	assert len(jk["position"]) == 260 and len(jk["shape"]) == 220 and len(jk["randoms_position"]) == 500
	assert set(np.unique(jk["randoms_position"])) == {0, 1, 2, 3}
	again = obj.assign_jackknife_patches(data, randoms, 4)
	assert all(np.array_equal(jk[k], again[k]) for k in jk)
	# nearest-centre property: a galaxy's patch is the patch of the centre (mean direction of its randoms) closest to it
	x = obj._unit_vectors(randoms["RA"], randoms["DEC"])
	centres = np.stack([x[jk["randoms_position"] == k].mean(axis=0) for k in range(4)])
	centres /= np.linalg.norm(centres, axis=1)[:, None]
	assert np.mean(obj._nearest_centre(obj._unit_vectors(data["RA"], data["DEC"]), centres) == jk["position"]) > 0.97
	# patches are compact on the sky: mean angular distance to the own centre well below the patch-to-patch distance
	own = np.degrees(np.arccos(np.clip(np.sum(x * centres[jk["randoms_position"]], axis=1), -1, 1))).mean()
	assert own < 1.5, own  # a 4 x 4 degree field in 4 patches
	monkeypatch.setattr(MeasureIALightcone, "_pair_sums", oracle_lc_pair_sums())
	with np.errstate(all="ignore"):
		obj.measure_xi_w("galaxies", "All", "both", num_jk=4)
	got = read_all(str(tmp_path / "k.hdf5"))
	assert got["w_g_plus/All_jackknife_cov_4"].shape == (5, 5) and "w/xi_gg/All_jk4/All_3_RR" in got
	assert sorted(obj.num_samples) == ["0", "1", "2", "3"] and obj.num_samples["0"]["D"] == int(np.sum(jk["position"] != 0))

"""Generate tests/golden/lc_*.npz from the UNMODIFIED reference's light-cone path.  TEST INFRASTRUCTURE ONLY; runs in the
build container (needs /root/reference), never on the GPU box.

    python oracle/make_golden_lightcone.py

Each fixture holds every dataset `measureia.MeasureIALightcone.measure_xi_w / measure_xi_multipoles` wrote for a small
seeded catalogue (measure_cov / calc_errors = False) plus the recipe to rebuild the inputs.  pyccl is replaced by
oracle/ref_shims/pyccl (same flat-LCDM distances as measure_ia_b200/cosmo.py): the fixtures pin everything downstream of
the distance conversion."""
import json
import os
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(_HERE), _HERE]

CONFIGS = {
	# name: (catalogue recipe, call)
	"lc_w_galaxies_both": (dict(n=500, n_shape=350, n_rand=900, seed=11, weights=True),
						   dict(kind="w", IA_estimator="galaxies", corr_type="both", over_h=False)),
	"lc_w_clusters_both_overh": (dict(n=400, n_shape=300, n_rand=700, seed=12, weights=False),
								 dict(kind="w", IA_estimator="clusters", corr_type="both", over_h=True)),
	"lc_w_clusters_gg_two_randoms": (dict(n=300, n_shape=300, n_rand=600, n_rand_shape=500, seed=13, weights=True),
									 dict(kind="w", IA_estimator="clusters", corr_type="gg", over_h=False)),
	"lc_w_galaxies_gplus_masked": (dict(n=450, n_shape=450, n_rand=450, seed=14, weights=True, masked=True),
								   dict(kind="w", IA_estimator="galaxies", corr_type="g+", over_h=False)),
	"lc_m_galaxies_both": (dict(n=500, n_shape=350, n_rand=900, seed=15, weights=True),
						   dict(kind="multipoles", IA_estimator="galaxies", corr_type="both", over_h=False)),
	"lc_m_clusters_both_overh_rpcut": (dict(n=400, n_shape=300, n_rand=700, seed=16, weights=False),
									   dict(kind="multipoles", IA_estimator="clusters", corr_type="both", over_h=True, rp_cut=1.5)),
	"lc_m_galaxies_gg": (dict(n=300, n_shape=300, n_rand=600, seed=17, weights=True),
						 dict(kind="multipoles", IA_estimator="galaxies", corr_type="gg", over_h=False)),
	# jackknife covariance with caller-supplied patches (RA stripes, labels 1..K); the reference needs num_nodes > 1 for it
	"lc_w_galaxies_both_jk4": (dict(n=420, n_shape=330, n_rand=700, seed=18, weights=True, jk=4),
							   dict(kind="w", IA_estimator="galaxies", corr_type="both", over_h=False)),
	"lc_w_clusters_gplus_jk3_two_randoms": (dict(n=360, n_shape=300, n_rand=600, n_rand_shape=520, seed=19, weights=True, jk=3),
											dict(kind="w", IA_estimator="clusters", corr_type="g+", over_h=False)),
	"lc_m_clusters_both_overh_jk3": (dict(n=380, n_shape=300, n_rand=640, seed=20, weights=False, jk=3),
									 dict(kind="multipoles", IA_estimator="clusters", corr_type="both", over_h=True)),
}
BINNING = dict(separation_limits=[0.5, 20.0], num_bins_r=5, num_bins_pi=6, pi_max=40.0)


def build_inputs(cat):
	"""Seeded light-cone catalogue: a 4 x 4 degree patch at 0.10 < z < 0.13 (about 12 x 12 x 120 Mpc)."""
	rng = np.random.default_rng(cat["seed"])

	def sky(k):
		return rng.uniform(10.0, 14.0, k), rng.uniform(-2.0, 2.0, k), rng.uniform(0.10, 0.13, k)

	n, ns = cat["n"], cat["n_shape"]
	ra, dec, z = sky(n)
	ra_s, dec_s, z_s = sky(ns)
	data = {"RA": ra, "DEC": dec, "Redshift": z, "RA_shape_sample": ra_s, "DEC_shape_sample": dec_s, "Redshift_shape_sample": z_s,
			"e1": rng.normal(0, 0.2, ns), "e2": rng.normal(0, 0.2, ns)}
	if cat.get("weights"):
		data["weight"], data["weight_shape_sample"] = rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, ns)
	r_ra, r_dec, r_z = sky(cat["n_rand"])
	randoms = {"RA": r_ra, "DEC": r_dec, "Redshift": r_z}
	if cat.get("n_rand_shape"):
		a, b, c = sky(cat["n_rand_shape"])
		randoms.update(RA_shape_sample=a, DEC_shape_sample=b, Redshift_shape_sample=c)
	masks = None
	if cat.get("masked"):  # the reference applies `masks` to the randoms-as-positions dictionaries too: equal lengths needed
		assert n == ns == cat["n_rand"]
		m = rng.random(n) < 0.75
		masks = {k: m.copy() for k in ("Redshift", "Redshift_shape_sample", "RA", "RA_shape_sample", "DEC", "DEC_shape_sample",
									   "e1", "e2", "weight", "weight_shape_sample")}
	return data, randoms, masks


def build_patches(cat, data, randoms):
	"""Jackknife patches as RA stripes, labels 1..K (k-means labels of the reference come from kmeans_radec, absent here)."""
	K = cat.get("jk")
	if not K:
		return None
	lab = lambda ra: np.minimum((np.asarray(ra) - 10.0) / 4.0 * K, K - 1).astype(int) + 1  # noqa: E731
	jk = {"position": lab(data["RA"]), "shape": lab(data["RA_shape_sample"])}
	if "RA_shape_sample" in randoms and randoms["RA_shape_sample"] is not randoms["RA"]:
		jk["randoms_position"], jk["randoms_shape"] = lab(randoms["RA"]), lab(randoms["RA_shape_sample"])
	else:
		jk["randoms"] = lab(randoms["RA"])
	return jk


def run_reference(cat, call):
	import run_reference as rr
	from measure_ia_b200 import h5lite
	measureia = rr.load_reference()
	data, randoms, masks = build_inputs(cat)
	tmp = tempfile.mkdtemp(prefix="mia_lc_")
	out = os.path.join(tmp, "out.hdf5")
	devnull, stdout = open(os.devnull, "w"), sys.stdout
	sys.stdout = devnull
	try:
		jk = build_patches(cat, data, randoms)
		obj = measureia.MeasureIALightcone(data, randoms, out, BINNING["separation_limits"], BINNING["num_bins_r"],
										   BINNING["num_bins_pi"], BINNING["pi_max"], 2 if jk else 1)
		kw = dict(masks=masks, over_h=call["over_h"], jk_patches=jk)
		with np.errstate(all="ignore"):
			if call["kind"] == "w":
				obj.measure_xi_w(call["IA_estimator"], "All", call["corr_type"], measure_cov=bool(jk), **kw)
			else:
				obj.measure_xi_multipoles(call["IA_estimator"], "All", call["corr_type"], calc_errors=bool(jk),
										  rp_cut=call.get("rp_cut"), **kw)
	finally:
		sys.stdout = stdout
	f = h5lite.File(out, "r")
	res = rr._flatten(f)
	f.close()
	return res


def main():
	golden = os.path.join(os.path.dirname(_HERE), "tests", "golden")
	for name, (cat, call) in CONFIGS.items():
		res = run_reference(cat, call)
		meta = dict(catalogue=cat, call=call, binning=BINNING)
		np.savez_compressed(os.path.join(golden, name + ".npz"), __meta__=json.dumps(meta),
							**{k.replace("/", "|"): v for k, v in res.items()})
		dd = [k for k in res if k.endswith("_DD") and "randoms" not in k]
		print(name, len(res), "datasets", {k: float(res[k].sum()) for k in dd})


if __name__ == "__main__":
	main()

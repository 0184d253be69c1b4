"""CPU ORACLE for the light-cone brute pair loops.  TEST INFRASTRUCTURE ONLY -- nothing in
measure_ia_b200/ imports this; only tests/, __graft_entry__.smoke() and bench.py's CPU leg may.

A numpy restatement, vectorised over blocks of position galaxies, of
    src/measureia/measure_w_lightcone.py:123-183   _measure_xi_rp_pi_lightcone_brute      (geom 'rppi', shapes=True)
    src/measureia/measure_w_lightcone.py:296-334   _count_pairs_xi_rp_pi_lightcone_brute  (geom 'rppi', shapes=False)
    src/measureia/measure_m_lightcone.py:116-191   _measure_xi_r_mur_lightcone_brute      (geom 'rmu',  shapes=True)
    src/measureia/measure_m_lightcone.py:286-330   _count_pairs_xi_r_mur_lightcone_brute  (geom 'rmu',  shapes=False)
with the reference's operation order and numpy calls (arctan2 / cos / sin / log10 are numpy's here as there, so on one host
the two agree to the last bit except for the order of the np.add.at sums).  Pinned by tests/golden/lc_*.npz, which
oracle/make_golden_lightcone.py generated from the UNMODIFIED reference (pyccl replaced by oracle/ref_shims/pyccl, i.e.
the distance conversion is shared with the product and NOT pinned against CCL).

Jackknife: `touch[k]` = sums over pairs with the position or the shape galaxy in patch k; the reference's realisation k
(both samples without patch k, measure_jackknife.py:116-134) equals total - touch[k] exactly for counts.
"""
import numpy as np


def distances(redshift, cosmology, over_h):
	"""measure_w_lightcone.py:123-133: chi = comoving_radial_distance(cosmo, 1 / (1 + z)), times h when over_h."""
	from measure_ia_b200.cosmo import Cosmology, comoving_radial_distance
	if cosmology is None:
		cosmology = Cosmology(Omega_c=0.225, Omega_b=0.045, sigma8=0.8, h=0.7, n_s=1.0)
	chi = comoving_radial_distance(cosmology, 1 / (1 + np.asarray(redshift, dtype=np.float64)))
	h = cosmology["h"]
	if over_h:
		chi = chi * h
	return chi, h


def shape_angles(e1, e2):
	"""measure_w_lightcone.py:135-140: e and the position angle of the (normalised) semi-major axis."""
	theta = 1. / 2 * np.arctan2(e2, e1)
	axis = np.array([np.cos(theta), np.sin(theta)])
	axis = axis / np.sqrt(np.sum(axis ** 2, axis=0))
	return np.sqrt(e1 ** 2 + e2 ** 2), np.arctan2(axis[1], axis[0])


def pair_sums(geom, pos, shp, r_min, r_max, r_bins, bins2, n_r, n_2, h=1.0, over_h=False, rp_cut=0.0, shapes=True,
			  patches_pos=None, patches_shape=None, num_patches=0, block=256):
	"""pos = dict(ra, dec, chi, w); shp = dict(ra, dec, chi, w[, e, phi]) (chi already times h when over_h).
	Returns dict(count, DD, SpD, ScD[, touch_count, touch_DD, touch_SpD])."""
	nb = n_r * n_2
	cnt = np.zeros(nb, dtype=np.int64)
	DD, SpD, ScD = np.zeros(nb), np.zeros(nb), np.zeros(nb)
	K = int(num_patches)
	t_cnt, t_DD, t_SpD = np.zeros((K, nb), dtype=np.int64), np.zeros((K, nb)), np.zeros((K, nb))
	d_logr = (np.log10(r_max) - np.log10(r_min)) / n_r
	d_2 = (bins2[-1] - bins2[0]) / n_2 if geom == "rppi" else 2.0 / n_2
	ra_s, dec_s, chi_s, w_s = shp["ra"], shp["dec"], shp["chi"], shp["w"]
	tested = 0
	for i0 in range(0, len(pos["ra"]), block):
		sl = slice(i0, min(i0 + block, len(pos["ra"])))
		ra_n, dec_n, chi_n, w_n = (pos[k][sl][:, None] for k in ("ra", "dec", "chi", "w"))
		los = chi_s[None, :] - chi_n
		dra = (ra_s[None, :] - ra_n) / 180 * np.pi
		ddec = (dec_s[None, :] - dec_n) / 180 * np.pi
		dx = dra * chi_n * np.cos(dec_n / 180 * np.pi)
		dy = ddec * chi_n
		px, py = (dx * h, dy * h) if over_h else (dx, dy)
		rp = np.sqrt(px ** 2 + py ** 2)
		tested += los.size
		if geom == "rppi":
			sep = rp
			mask = (sep >= r_bins[0]) * (sep < r_bins[-1]) * (los >= bins2[0]) * (los < bins2[-1])
			second = los
		else:
			sep = np.sqrt((dx ** 2 + dy ** 2) + los ** 2)
			mask = (rp > rp_cut) * (sep >= r_bins[0]) * (sep < r_bins[-1])
			with np.errstate(invalid="ignore", divide="ignore"):
				second = los / sep
		ip, js = np.nonzero(mask)
		if len(ip) == 0:
			continue
		ind_r = np.floor(np.log10(sep[ip, js]) / d_logr - np.log10(r_bins[0]) / d_logr).astype(int)
		ind_2 = np.floor(second[ip, js] / d_2 - bins2[0] / d_2).astype(int)
		ind_r[ind_r == n_r] = n_r - 1  # measure_w_lightcone.py:173-176 (the count / multipole loops would raise instead)
		ind_2[ind_2 == n_2] = n_2 - 1
		b = ind_r * n_2 + ind_2
		ww = (w_n * w_s[None, :])[ip, js]
		np.add.at(cnt, b, 1)
		np.add.at(DD, b, ww)
		tp = None
		if shapes:
			with np.errstate(invalid="ignore", divide="ignore"):
				phi_sep = np.arctan2((py / rp)[ip, js], (px / rp)[ip, js])
			phi = shp["phi"][js] - phi_sep
			e_plus, e_cross = -shp["e"][js] * np.cos(2 * phi), -shp["e"][js] * np.sin(2 * phi)
			e_plus[np.isnan(e_plus)] = 0.0
			e_cross[np.isnan(e_cross)] = 0.0
			tp = ww * e_plus
			np.add.at(SpD, b, tp)
			np.add.at(ScD, b, ww * e_cross)
		if K:
			pp, ps = patches_pos[sl][ip], patches_shape[js]
			for lab, sel in ((ps, slice(None)), (pp, pp != ps)):
				np.add.at(t_cnt, (lab[sel], b[sel]), 1)
				np.add.at(t_DD, (lab[sel], b[sel]), ww[sel])
				if shapes:
					np.add.at(t_SpD, (lab[sel], b[sel]), tp[sel])
	sh = (n_r, n_2)
	out = dict(count=cnt.reshape(sh), DD=DD.reshape(sh), SpD=SpD.reshape(sh), ScD=ScD.reshape(sh), tested=tested)
	if K:
		out.update(touch_count=t_cnt.reshape((K,) + sh), touch_DD=t_DD.reshape((K,) + sh), touch_SpD=t_SpD.reshape((K,) + sh))
	return out


def sample(ra, dec, redshift, weight=None, cosmology=None, over_h=False, e1=None, e2=None):
	"""Per-galaxy preparation of one catalogue for `pair_sums` (measure_w_lightcone.py:80-140)."""
	chi, h = distances(redshift, cosmology, over_h)
	d = dict(ra=np.asarray(ra, dtype=np.float64), dec=np.asarray(dec, dtype=np.float64), chi=chi,
			 w=np.ones(len(chi)) if weight is None else np.asarray(weight, dtype=np.float64))
	if e1 is not None:
		d["e"], d["phi"] = shape_angles(np.asarray(e1, dtype=np.float64), np.asarray(e2, dtype=np.float64))
	return d, h

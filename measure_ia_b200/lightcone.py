"""Host-side mirror of the reference's LIGHT-CONE API on top of the B200 brute pair-loop operator.

``MeasureIALightcone`` keeps the constructor, the ``measure_xi_w`` / ``measure_xi_multipoles`` signatures, the data dictionaries
(``RA``, ``DEC``, ``Redshift``, ``e1``, ``e2``, weights; randoms), the estimators and the HDF5 layout of the reference
(``src/measureia/measure_IA.py:265-1058``, ``measure_w_lightcone.py``, ``measure_m_lightcone.py``,
``measure_IA_base.py:670-742``).  The four O(N_p N_s) Python loops over position galaxies
(``_measure_xi_rp_pi_lightcone_brute``, ``_count_pairs_xi_rp_pi_lightcone_brute`` and their (r, mu_r) twins) are ONE operator,
``ops.lightcone_paircount`` (CUDA, sm_100a; no CPU fallback).

Jackknife covariance (``measure_cov`` / ``calc_errors``) with caller-supplied ``jk_patches``: the reference re-runs every loop K
times without patch k (measure_jackknife.py:265-483, one process per patch); here the operator returns, in the SAME pass as
the totals, the sums over pairs touching each patch, and realisation k = total - touch[k].  The files follow the reference's
multiprocessing branch (its single-process branch, measure_jackknife.py:307-310, passes its arguments in the wrong order and
cannot run).  ``num_jk`` without ``jk_patches``: the reference clusters the randoms on the sky with ``kmeans_radec``
(measure_IA_base.py:744-803), which is not in this image; ``assign_jackknife_patches`` uses it when importable and otherwise a
spherical k-means of its own (same role: centres from the position randoms, nearest centre for the other samples; the labels are
a different -- equally arbitrary -- partition, since kmeans_radec starts from unseeded random centres).
"""
from __future__ import annotations

import time

import numpy as np

from . import cosmo
from .box import MeasureIABase
from .io import create_group_hdf5, open_file, write_dataset_hdf5
from .jackknife import JackknifeCombinationMixin


class MeasureIALightcone(MeasureIABase, JackknifeCombinationMixin):
	"""Drop-in for ``measureia.MeasureIALightcone`` (measure_IA.py:265-1058); jackknife patches must be supplied (``jk_patches``)."""

	def __init__(self, data, randoms_data, output_file_name, separation_limits=[0.1, 20.0], num_bins_r=8, num_bins_pi=20,
				 pi_max=None, num_nodes=1):
		super().__init__(data, output_file_name, False, None, separation_limits, num_bins_r, num_bins_pi, pi_max, None, False)
		self.num_nodes = num_nodes  # accepted for compatibility
		self.randoms_data = randoms_data
		self.data_dir = None
		self.num_samples = None
		self.device = None
		self.last_stats = None

	# ---- the pair loop ---------------------------------------------------------------------------------------------------
	def _device(self):
		import torch
		if not torch.cuda.is_available():
			raise RuntimeError("measure_ia_b200 needs a CUDA device: the light-cone pair operator has no CPU fallback")
		return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

	def _select(self, masks, shapes):
		"""The reference's input selection (measure_w_lightcone.py:80-112): arrays of ``self.data``, masked key by key; missing
		weight masks become all-True masks stored in the caller's dict."""
		d = self.data
		keys = ["Redshift", "Redshift_shape_sample", "RA", "RA_shape_sample", "DEC", "DEC_shape_sample"] + (["e1", "e2"] if shapes else [])
		if masks is None:
			out = {k: d[k] for k in keys}
			out["weight"], out["weight_shape_sample"] = d["weight"], d["weight_shape_sample"]
			return out
		out = {k: d[k][masks[k]] for k in keys}
		if "weight" not in masks:
			masks["weight"] = np.ones(self.Num_position, dtype=bool)
		if "weight_shape_sample" not in masks:
			masks["weight_shape_sample"] = np.ones(self.Num_shape, dtype=bool)
		out["weight"] = d["weight"][masks["weight"]]
		out["weight_shape_sample"] = d["weight_shape_sample"][masks["weight_shape_sample"]]
		return out

	def _pair_sums(self, geom, shapes, masks, over_h, cosmology, rp_cut=None, patches=None):
		"""Per-galaxy preparation (distances, cos(dec), shape angles: O(N)), then ONE operator call for the O(N_p N_s) loop.
		Returns dict(count, DD, SpD, ScD[, touch_*])."""
		import torch

		from . import ops
		t0 = time.perf_counter()
		dev = self._device()
		sel = self._select(masks, shapes)
		if cosmology is None:  # measure_w_lightcone.py:123-126
			cosmology = cosmo.Cosmology(Omega_c=0.225, Omega_b=0.045, sigma8=0.8, h=0.7, n_s=1.0)
		h = cosmology["h"]
		f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
		up = lambda a: torch.from_numpy(f(a)).to(dev)  # noqa: E731
		# cos(dec) decides which bin a pair falls in and numpy's cos differs from CUDA's in the last bit: it stays numpy's.
		# Everything else per galaxy is +, -, *, /, sqrt (IEEE-exact on the device: the same bits as numpy) or only enters the
		# sums (shape angles), so it runs on the device in fp64 once the raw columns are uploaded.
		pos = dict(ra=up(sel["RA"]), dec=up(sel["DEC"]), cosdec=up(np.cos(f(sel["DEC"]) / 180 * np.pi)), weight=up(sel["weight"]))
		shp = dict(ra=up(sel["RA_shape_sample"]), dec=up(sel["DEC_shape_sample"]), weight=up(sel["weight_shape_sample"]))
		if cosmo.is_builtin_flat_lcdm(cosmology):
			om = cosmology["Omega_c"] + cosmology["Omega_b"]
			pos["chi"] = cosmo.flat_lcdm_distance(om, h, 1 / (1 + up(sel["Redshift"])))
			shp["chi"] = cosmo.flat_lcdm_distance(om, h, 1 / (1 + up(sel["Redshift_shape_sample"])))
		else:  # pyccl, or a caller-supplied chi(a): on the host
			pos["chi"] = up(cosmo.comoving_radial_distance(cosmology, 1 / (1 + f(sel["Redshift"]))))
			shp["chi"] = up(cosmo.comoving_radial_distance(cosmology, 1 / (1 + f(sel["Redshift_shape_sample"]))))
		if over_h:  # :131-133
			pos["chi"], shp["chi"] = pos["chi"] * h, shp["chi"] * h
		if shapes:  # :135-140; e cos 2phi_axis and e sin 2phi_axis are what the operator needs
			e1, e2 = up(sel["e1"]), up(sel["e2"])
			theta = 1. / 2 * torch.atan2(e2, e1)
			a0, a1 = torch.cos(theta), torch.sin(theta)
			norm = torch.sqrt(a0 ** 2 + a1 ** 2)
			phi_axis = torch.atan2(a1 / norm, a0 / norm)
			e = torch.sqrt(e1 ** 2 + e2 ** 2)
			shp["e1"], shp["e2"] = e * torch.cos(2 * phi_axis), e * torch.sin(2 * phi_axis)
		num_patches = 0
		if patches is not None:
			pp, ps = np.asarray(patches[0]), np.asarray(patches[1])
			if masks is not None and len(pp) == len(masks["Redshift"]) and len(ps) == len(masks["Redshift_shape_sample"]):
				pp, ps = pp[masks["Redshift"]], ps[masks["Redshift_shape_sample"]]  # labels of the full samples (measure_jackknife.py:351-355)
			lo = int(min(pp.min(), ps.min())) if len(pp) and len(ps) else 0
			num_patches = (int(max(pp.max(), ps.max())) - lo + 1) if len(pp) and len(ps) else 0
			pos["patch"] = torch.from_numpy(np.ascontiguousarray((pp - lo).astype(np.int32))).to(dev)
			shp["patch"] = torch.from_numpy(np.ascontiguousarray((ps - lo).astype(np.int32))).to(dev)
		for k, dd in (("position", pos), ("shape", shp)):
			if any(v.shape[0] != dd["ra"].shape[0] for v in dd.values()):
				raise ValueError(f"light-cone {k} sample: arrays of different lengths")

		def upload(d):  # sorted by chi -- the operator's one requirement (it culls on the chi window)
			order = torch.argsort(d["chi"], stable=True)
			return {k: v[order].contiguous() for k, v in d.items()}

		r2_thr, thr2, rp2_cut, clean = self._thresholds_for("rppi" if geom == "rppi" else "rmu", rp_cut)
		rank, world = 0, 1
		if torch.distributed.is_available() and torch.distributed.is_initialized():
			rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
		t1 = time.perf_counter()
		out = ops.lightcone_paircount(upload(pos), upload(shp), torch.from_numpy(r2_thr), torch.from_numpy(thr2),
									  ops.GEOM_RPPI if geom == "rppi" else ops.GEOM_RMU, bool(shapes), num_patches,
									  float(h) if over_h else 1.0, float(rp2_cut), rank, world)
		if world > 1:
			from .box import combine_across_ranks
			out = list(combine_across_ranks(out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]))
		torch.cuda.synchronize(dev)
		t2 = time.perf_counter()
		cnt, ddw, spd, scd, t_cnt, t_w, t_spd, stats = (t.cpu().numpy() for t in out)
		self.last_stats = dict(tested=int(stats[0]), binned=int(stats[1]), kernel=int(stats[4]), launches=int(stats[7]),
							   thresholds_clean=bool(clean), rank=rank, world=world, t_prep=t1 - t0, t_device=t2 - t1,
							   kernel_ms=ops.LAST_LC_TIMINGS_MS[0])
		res = dict(count=cnt, DD=ddw, SpD=spd, ScD=scd)
		if num_patches:
			res.update(touch_count=t_cnt, touch_DD=t_w, touch_SpD=t_spd, patch_lo=lo, patches=(pp, ps))
		self.last_result = res
		return res

	# ---- the reference's four loops, same names, same outputs ------------------------------------------------------------------
	def _centres(self, geom):
		sep = self.r_bins[:-1] + abs((self.r_bins[1:] - self.r_bins[:-1]) / 2.0)
		b2 = self.pi_bins if geom == "rppi" else self.mu_r_bins
		return sep, b2[:-1] + abs((b2[1:] - b2[:-1]) / 2.0)

	def _write_realisations(self, geom, res, dataset_name, shapes, data_suffix, sample_names):
		"""Leave-one-patch-out realisations from the per-patch sums of the SAME operator call, in the layout of the reference's
		multiprocessing branch (measure_jackknife.py:433-474): `xi_g_plus/<X>_jk<K>/<X>_<i>_SplusD`, `xi_gg/<X>_jk<K>/<X>_<i>_DD`
		(or `<X>_<i><suffix>` for the pair counts), each with its bin centres; fills `self.num_samples[str(i)]` (:356-357)."""
		pp, ps = res["patches"]
		lo_p, hi_p = int(min(pp)), int(max(pp))  # group name and loop range follow the POSITION sample's labels (:337-338)
		K = hi_p - lo_p + 1
		sep, mid2 = self._centres(geom)
		top = "w" if geom == "rppi" else "multipoles"
		n1, n2 = ("_rp", "_pi") if geom == "rppi" else ("_r", "_mu_r")
		writer = self.last_stats["rank"] == 0 and self.output_file_name is not None
		f = open_file(self.output_file_name, "a") if writer else None
		try:
			for i in range(lo_p, hi_p + 1):
				self.num_samples.setdefault(f"{i}", {})
				self.num_samples[f"{i}"][sample_names[0]] = int(np.sum(ps != i))
				self.num_samples[f"{i}"][sample_names[1]] = int(np.sum(pp != i))
				if not writer:
					continue
				k = i - res["patch_lo"]
				left = res["count"] - res["touch_count"][k]
				DD = np.where(left > 0, res["DD"] - res["touch_DD"][k], 0.0)  # exactly 0 where no pair is left (no rounding residue)
				DD[np.where(DD == 0)] = 1  # measure_w_lightcone.py:188
				if shapes:
					SpD = np.where(left > 0, res["SpD"] - res["touch_SpD"][k], 0.0)
					g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_g_plus/{dataset_name}_jk{K}")
					write_dataset_hdf5(g, f"{dataset_name}_{i}_SplusD", data=SpD)
					write_dataset_hdf5(g, f"{dataset_name}_{i}{n1}", data=sep)
					write_dataset_hdf5(g, f"{dataset_name}_{i}{n2}", data=mid2)
				g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_gg/{dataset_name}_jk{K}")
				write_dataset_hdf5(g, f"{dataset_name}_{i}" + ("_DD" if shapes else data_suffix), data=DD)
				write_dataset_hdf5(g, f"{dataset_name}_{i}{n1}", data=sep)
				write_dataset_hdf5(g, f"{dataset_name}_{i}{n2}", data=mid2)
		finally:
			if f is not None:
				f.close()

	def _measure_brute(self, geom, dataset_name, masks, return_output, print_num, over_h, cosmology, jk_group_name, rp_cut=None,
					   jk=None):
		res = self._pair_sums(geom, True, masks, over_h, cosmology, rp_cut, patches=None if jk is None else jk[:2])
		if jk is not None:
			self._write_realisations(geom, res, dataset_name, True, "_DD", jk[2])
		if print_num and self.verbose:
			print(f"There are {len(self.data['RA_shape_sample'])} galaxies in the shape sample and {len(self.data['RA'])} galaxies in the position sample.")
		DD, SpD, ScD = res["DD"].copy(), res["SpD"], res["ScD"]
		DD[np.where(DD == 0)] = 1  # measure_w_lightcone.py:188
		correlation = SpD / DD
		sep, mid2 = self._centres(geom)
		top = "w" if geom == "rppi" else "multipoles"
		n1, n2 = ("_rp", "_pi") if geom == "rppi" else ("_r", "_mu_r")
		if self.output_file_name is not None and not return_output:
			if self.last_stats["rank"] == 0:
				f = open_file(self.output_file_name, "a")
				try:
					g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_g_plus/{jk_group_name}")
					write_dataset_hdf5(g, dataset_name, data=correlation)
					write_dataset_hdf5(g, dataset_name + "_SplusD", data=SpD)
					write_dataset_hdf5(g, dataset_name + n1, data=sep)
					write_dataset_hdf5(g, dataset_name + n2, data=mid2)
					g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_g_cross/{jk_group_name}")
					write_dataset_hdf5(g, dataset_name + "_ScrossD", data=ScD)
					write_dataset_hdf5(g, dataset_name + n1, data=sep)
					write_dataset_hdf5(g, dataset_name + n2, data=mid2)
					g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_gg/{jk_group_name}")
					write_dataset_hdf5(g, dataset_name + "_DD", data=DD)
					write_dataset_hdf5(g, dataset_name + n1, data=sep)
					write_dataset_hdf5(g, dataset_name + n2, data=mid2)
				finally:
					f.close()
			return None
		return SpD, DD, sep, mid2

	def _count_brute(self, geom, dataset_name, masks, return_output, print_num, over_h, cosmology, data_suffix, jk_group_name,
					 rp_cut=None, jk=None):
		res = self._pair_sums(geom, False, masks, over_h, cosmology, rp_cut, patches=None if jk is None else jk[:2])
		if jk is not None:
			self._write_realisations(geom, res, dataset_name, False, data_suffix, jk[2])
		DD = res["DD"].copy()
		DD[np.where(DD == 0)] = 1  # measure_w_lightcone.py:339
		sep, mid2 = self._centres(geom)
		top = "w" if geom == "rppi" else "multipoles"
		n1, n2 = ("_rp", "_pi") if geom == "rppi" else ("_r", "_mu_r")
		if self.output_file_name is not None and not return_output:
			if self.last_stats["rank"] == 0:
				f = open_file(self.output_file_name, "a")
				try:
					g = create_group_hdf5(f, f"{self.snap_group}/{top}/xi_gg/{jk_group_name}")
					write_dataset_hdf5(g, dataset_name + data_suffix, data=DD)
					write_dataset_hdf5(g, dataset_name + n1, data=sep)
					write_dataset_hdf5(g, dataset_name + n2, data=mid2)
				finally:
					f.close()
			return None
		return DD, sep, mid2

	def _measure_xi_rp_pi_lightcone_brute(self, dataset_name, masks=None, return_output=False, print_num=True, over_h=False,
										  cosmology=None, jk_group_name="", jk=None):
		"""measure_w_lightcone.py:45-214.  `jk` (not in the reference): (patches_position, patches_shape, num_sample_names) --
		also write the leave-one-patch-out realisations from the same operator call."""
		return self._measure_brute("rppi", dataset_name, masks, return_output, print_num, over_h, cosmology, jk_group_name, jk=jk)

	def _count_pairs_xi_rp_pi_lightcone_brute(self, dataset_name, masks=None, return_output=False, print_num=True, over_h=False,
											  cosmology=None, data_suffix="_DD", jk_group_name="", jk=None):
		"""measure_w_lightcone.py:216-343."""
		return self._count_brute("rppi", dataset_name, masks, return_output, print_num, over_h, cosmology, data_suffix, jk_group_name,
								 jk=jk)

	def _measure_xi_r_mur_lightcone_brute(self, dataset_name, masks=None, return_output=False, print_num=True, over_h=True,
										  cosmology=None, rp_cut=None, jk_group_name="", jk=None):
		"""measure_m_lightcone.py:45-219."""
		return self._measure_brute("rmu", dataset_name, masks, return_output, print_num, over_h, cosmology, jk_group_name, rp_cut,
								   jk=jk)

	def _count_pairs_xi_r_mur_lightcone_brute(self, dataset_name, masks=None, return_output=False, print_num=True, over_h=False,
											  cosmology=None, rp_cut=None, data_suffix="_DD", jk_group_name="", jk=None):
		"""measure_m_lightcone.py:221-363."""
		return self._count_brute("rmu", dataset_name, masks, return_output, print_num, over_h, cosmology, data_suffix, jk_group_name,
								 rp_cut, jk=jk)

	# ---- estimators (measure_IA_base.py:670-742) -----------------------------------------------------------------------------
	def _obs_estimator(self, corr_type, IA_estimator, dataset_name, dataset_name_randoms, num_samples, jk_group_name="",
					   jk_group_name_randoms=""):
		if IA_estimator not in ("clusters", "galaxies"):
			raise ValueError("Unknown input for IA_estimator, choose from [clusters, galaxies].")
		f = open_file(self.output_file_name, "a")
		try:
			which, top = corr_type
			gp = which in ("g+", "both")
			gg = which in ("gg", "both")
			if gp:
				group_gp = f[f"{self.snap_group}/{top}/xi_g_plus/{jk_group_name}"]
				group_gp_r = f[f"{self.snap_group}/{top}/xi_g_plus/{jk_group_name_randoms}"]
				SpD = group_gp[f"{dataset_name}_SplusD"][:]
				SpR = group_gp_r[f"{dataset_name_randoms}_SplusD"][:]
			group_gg = f[f"{self.snap_group}/{top}/xi_gg/{jk_group_name}"]
			group_gg_r = f[f"{self.snap_group}/{top}/xi_gg/{jk_group_name_randoms}"]
			DD = group_gg[f"{dataset_name}_DD"][:]
			fD = num_samples["D"] / num_samples["R_D"]
			fS = num_samples["S"] / num_samples["R_S"] if (gg or IA_estimator == "galaxies") else None  # (not set for clusters / g+)
			read_SR = lambda: (group_gg[f"{dataset_name}_SR"][:] if which == "gg" else group_gg_r[f"{dataset_name_randoms}_DD"][:])  # noqa: E731
			with np.errstate(divide="ignore", invalid="ignore"):
				if IA_estimator == "clusters":
					SR = read_SR()
					SR *= fD
					if gp:
						SpR *= fD
						write_dataset_hdf5(group_gp, dataset_name, SpD / DD - SpR / SR)
					if gg:
						RD = group_gg[f"{dataset_name}_RD"][:]
						RR = group_gg[f"{dataset_name}_RR"][:]
						RD *= fS
						RR *= fS * fD
						write_dataset_hdf5(group_gg, dataset_name, (DD - RD - SR) / RR - 1)
				else:
					RR = group_gg[f"{dataset_name}_RR"][:]
					RR *= fS * fD
					if gp:
						SpR *= fD
						write_dataset_hdf5(group_gp, dataset_name, (SpD - SpR) / RR)
					if gg:
						RD = group_gg[f"{dataset_name}_RD"][:]
						SR = read_SR()
						RD *= fS
						SR *= fD
						write_dataset_hdf5(group_gg, dataset_name, (DD - RD - SR) / RR + 1)
		finally:
			f.close()

	def _measure_jackknife_covariance_lightcone(self, IA_estimator, corr_type, dataset_name, max_patch, min_patch=1,
												randoms_suf="_randoms"):
		"""Estimators and integrated statistics per realisation, then mean / std / covariance (measure_jackknife.py:172-263)."""
		try:
			kinds = {"both": ["_g_plus", "_gg"], "g+": ["_g_plus"], "gg": ["_gg"]}[corr_type[0]]
		except KeyError:
			raise KeyError("Unknown value for corr_type. Choose from [g+, gg, both]")
		K = max_patch - min_patch + 1
		covs, stds = [], []
		for kind in kinds:
			for b in range(min_patch, max_patch + 1):
				self._obs_estimator(corr_type, IA_estimator, f"{dataset_name}_{b}", f"{dataset_name}{randoms_suf}_{b}",
									self.num_samples[f"{b}"], jk_group_name=f"{dataset_name}_jk{K}",
									jk_group_name_randoms=f"{dataset_name}{randoms_suf}_jk{K}")
				if corr_type[1] == "w":
					self._measure_w_g_i(corr_type=corr_type[0], dataset_name=f"{dataset_name}_{b}", jk_group_name=f"{dataset_name}_jk{K}")
				else:
					self._measure_multipoles(corr_type=corr_type[0], dataset_name=f"{dataset_name}_{b}",
											 jk_group_name=f"{dataset_name}_jk{K}")
			f = open_file(self.output_file_name, "a")
			try:
				grp = f[f"{self.snap_group}/{corr_type[1]}{kind}/{dataset_name}_jk{K}"]
				reals = np.array([grp[f"{dataset_name}_{b}"][:] for b in range(min_patch, max_patch + 1)])
				with np.errstate(invalid="ignore"):
					mean, std, cov = self._jackknife_stats(reals)
				out = create_group_hdf5(f, f"{self.snap_group}/{corr_type[1]}{kind}")
				write_dataset_hdf5(out, f"{dataset_name}_mean_{K}", data=mean)
				write_dataset_hdf5(out, f"{dataset_name}_jackknife_{K}", data=std)
				write_dataset_hdf5(out, f"{dataset_name}_jackknife_cov_{K}", data=cov)
			finally:
				f.close()
			covs.append(cov)
			stds.append(std)
		return covs, stds

	# ---- jackknife patches on the sky (measure_IA_base.py:744-803) ---------------------------------------------------------------
	@staticmethod
	def _unit_vectors(ra, dec):
		ra, dec = np.radians(np.asarray(ra, dtype=np.float64)), np.radians(np.asarray(dec, dtype=np.float64))
		return np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)], axis=1)

	@staticmethod
	def _nearest_centre(x, centres, block=1 << 18):
		out = np.empty(len(x), dtype=np.int64)
		for i in range(0, len(x), block):
			out[i:i + block] = np.argmax(x[i:i + block] @ centres.T, axis=1)  # largest cosine = smallest angle
		return out

	def assign_jackknife_patches(self, data, randoms_data, num_jk, maxiter=100, tol=1.0e-5, seed=0):
		"""Patch labels 0..num_jk-1 for the four samples: k-means of the position randoms on the sphere, nearest centre for
		the shape randoms and the data (the reference: ``kmeans_radec.kmeans_sample(..., maxiter=100, tol=1.0e-5)`` then
		``find_nearest``).  Uses kmeans_radec when it is installed, else Lloyd iterations on unit vectors from `num_jk`
		randoms drawn with ``default_rng(seed)`` (reproducible; kmeans_radec's own start is unseeded)."""
		try:  # pragma: no cover - kmeans_radec is absent from the build image
			from kmeans_radec import kmeans_sample
			km = kmeans_sample(np.column_stack((randoms_data["RA"], randoms_data["DEC"])), num_jk, maxiter=maxiter, tol=tol)
			near = lambda ra, dec: km.find_nearest(np.column_stack((ra, dec)))  # noqa: E731
			first = km.labels
		except ImportError:
			x = self._unit_vectors(randoms_data["RA"], randoms_data["DEC"])
			if num_jk < 1 or num_jk > len(x):
				raise ValueError("num_jk must be between 1 and the number of randoms")
			centres = x[np.random.default_rng(seed).choice(len(x), size=num_jk, replace=False)].copy()
			for _ in range(maxiter):
				lab = self._nearest_centre(x, centres)
				new = np.zeros_like(centres)
				np.add.at(new, lab, x)
				norm = np.sqrt(np.sum(new ** 2, axis=1))
				empty = norm == 0.0
				new[~empty] /= norm[~empty, None]
				new[empty] = centres[empty]  # a centre that lost all its points stays where it is
				shift = float(np.max(np.arccos(np.clip(np.sum(new * centres, axis=1), -1.0, 1.0))))
				centres = new
				if shift < np.radians(tol):
					break
			near = lambda ra, dec: self._nearest_centre(self._unit_vectors(ra, dec), centres)  # noqa: E731
			first = near(randoms_data["RA"], randoms_data["DEC"])
		return {"randoms_position": first,
				"randoms_shape": near(randoms_data["RA_shape_sample"], randoms_data["DEC_shape_sample"]),
				"position": near(data["RA"], data["DEC"]),
				"shape": near(data["RA_shape_sample"], data["DEC_shape_sample"])}

	# ---- public API -------------------------------------------------------------------------------------------------------------
	def _prepare_randoms(self):
		"""measure_IA.py:396-413: one random sample serves both roles; default unit weights."""
		r = self.randoms_data
		one = "RA_shape_sample" not in r
		if one:
			r["RA_shape_sample"], r["DEC_shape_sample"], r["Redshift_shape_sample"] = r["RA"], r["DEC"], r["Redshift"]
		if "weight" not in r:
			r["weight"] = np.ones(len(r["RA"]))
		if "weight_shape_sample" not in r:
			r["weight_shape_sample"] = r["weight"] if one else np.ones(len(r["RA_shape_sample"]))
		return one

	def _measure(self, top, IA_estimator, dataset_name, corr_type, want_cov, masks, masks_randoms, cosmology, over_h, rp_cut,
				 jk_patches=None, num_jk=None):
		if IA_estimator not in ("clusters", "galaxies"):
			raise KeyError("Unknown input for IA_estimator, choose from [clusters, galaxies].")
		if corr_type not in ("both", "g+", "gg"):
			raise KeyError("Unknown value for corr_type. Choose from [g+, gg, both]")
		if want_cov:
			if jk_patches is None and num_jk is None:
				raise ValueError("Set calc_errors to False, or provide either jk_patches or num_jk input.")
			if corr_type == "gg":
				# the reference's estimator opens the `<name>_randoms_jk<K>` group it only writes for g+ / both
				# (measure_IA_base.py:703) and fails with this KeyError
				raise KeyError(f"Unable to open object (object '{dataset_name}_randoms_jk' doesn't exist): the light-cone jackknife "
							   "needs corr_type 'g+' or 'both'")
		geom = "rppi" if top == "w" else "rmu"
		kw = dict(over_h=over_h, cosmology=cosmology)
		if geom == "rmu":
			kw["rp_cut"] = rp_cut
		measure = self._measure_xi_rp_pi_lightcone_brute if geom == "rppi" else self._measure_xi_r_mur_lightcone_brute
		count = self._count_pairs_xi_rp_pi_lightcone_brute if geom == "rppi" else self._count_pairs_xi_r_mur_lightcone_brute
		data = self.data  # restored at the end (measure_IA.py:392,687)
		one_random_sample = self._prepare_randoms()
		if want_cov and jk_patches is None:  # measure_IA.py:415-418
			jk_patches = self.assign_jackknife_patches(data, self.randoms_data, num_jk)
		if want_cov and one_random_sample and "randoms" in jk_patches:  # measure_IA.py:420-423
			jk_patches["randoms_position"] = jk_patches["randoms"]
			jk_patches["randoms_shape"] = jk_patches["randoms"]
		self.data_dir = D = data
		if "weight" not in D:
			D["weight"] = np.ones(len(D["RA"]))
		if "weight_shape_sample" not in D:
			D["weight_shape_sample"] = np.ones(len(D["RA_shape_sample"]))
		R = self.randoms_data
		n = {}
		n["D"] = len(D["RA"]) if masks is None else len(D["RA"][masks["RA"]])
		n["S"] = len(D["RA_shape_sample"]) if masks is None else len(D["RA_shape_sample"][masks["RA_shape_sample"]])
		n["R_D"] = len(R["RA"]) if masks_randoms is None else len(R["RA"][masks_randoms["RA"]])
		n["R_S"] = len(R["RA_shape_sample"]) if masks_randoms is None else len(R["RA_shape_sample"][masks_randoms["RA_shape_sample"]])
		self.num_samples = n
		jk = (lambda a, b, names: None) if not want_cov else (lambda a, b, names: (jk_patches[a], jk_patches[b], names))
		if want_cov:
			self.num_samples = {}  # per patch from here on (measure_IA.py:551-554); `n` stays the totals' dictionary

		def combo(position, shape, shapes):
			"""The reference's temporary data dictionaries (measure_IA.py:460-552): position sample from `position`,
			shape sample from `shape`."""
			d = {"Redshift": position["Redshift"], "Redshift_shape_sample": shape["Redshift_shape_sample"], "RA": position["RA"],
				 "RA_shape_sample": shape["RA_shape_sample"], "DEC": position["DEC"], "DEC_shape_sample": shape["DEC_shape_sample"],
				 "weight": position["weight"], "weight_shape_sample": shape["weight_shape_sample"]}
			if shapes:
				d["e1"], d["e2"] = shape["e1"], shape["e2"]
			return d

		try:
			if corr_type in ("g+", "both"):
				self.data = D  # S+D
				measure(masks=masks, dataset_name=dataset_name, jk=jk("position", "shape", ["S", "D"]), **kw)
				self.data = combo(R, D, True)  # S+R
				measure(masks=masks, dataset_name=f"{dataset_name}_randoms", jk=jk("randoms_position", "shape", ["S", "R_D"]), **kw)
			if corr_type == "gg":  # SD, SR (already there for 'both')
				self.data = combo(D, D, False)
				count(masks=masks, dataset_name=dataset_name, data_suffix="_DD", **kw)
				self.data = combo(R, D, False)
				count(masks=masks, dataset_name=dataset_name, data_suffix="_SR", **kw)
			if corr_type in ("gg", "both"):  # RD
				self.data = combo(D, R, False)
				count(masks=masks, dataset_name=dataset_name, data_suffix="_RD", jk=jk("position", "randoms_shape", ["R_S", "D"]), **kw)
			if IA_estimator == "galaxies" or corr_type in ("gg", "both"):  # RR
				self.data = combo(R, R, False)
				count(masks=masks, dataset_name=dataset_name, data_suffix="_RR",
					  jk=jk("randoms_position", "randoms_shape", ["R_S", "R_D"]), **kw)
			if self.last_stats["rank"] == 0:
				self._obs_estimator([corr_type, top], IA_estimator, dataset_name, f"{dataset_name}_randoms", n)
				if top == "w":
					self._measure_w_g_i(corr_type=corr_type, dataset_name=dataset_name, return_output=False)
				else:
					self._measure_multipoles(corr_type=corr_type, dataset_name=dataset_name, return_output=False)
				if want_cov:
					self._measure_jackknife_covariance_lightcone(IA_estimator, [corr_type, top], dataset_name,
																 max_patch=int(max(jk_patches["shape"])),
																 min_patch=int(min(jk_patches["shape"])), randoms_suf="_randoms")
		finally:
			self.data = data

	def measure_xi_w(self, IA_estimator, dataset_name, corr_type, jk_patches=None, num_jk=None, measure_cov=True, masks=None,
					 masks_randoms=None, cosmology=None, over_h=False):
		"""xi_gg, xi_g+ and w_gg, w_g+ for light-cone data (measure_IA.py:336-688)."""
		self._measure("w", IA_estimator, dataset_name, corr_type, measure_cov, masks, masks_randoms, cosmology, over_h, None,
					  jk_patches, num_jk)

	def measure_xi_multipoles(self, IA_estimator, dataset_name, corr_type, jk_patches=None, num_jk=None, calc_errors=True,
							  masks=None, masks_randoms=None, cosmology=None, over_h=False, rp_cut=None):
		"""Multipoles for light-cone data (measure_IA.py:690-1058)."""
		self._measure("multipoles", IA_estimator, dataset_name, corr_type, calc_errors, masks, masks_randoms, cosmology, over_h,
					  rp_cut, jk_patches, num_jk)

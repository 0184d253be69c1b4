"""Combination of jackknife realisations of several datasets into (block) covariance matrices.

Mirrors reference ``src/measureia/measure_jackknife.py:485-648`` (``MeasureJackknife.measure_covariance_multiple_datasets``,
``create_full_cov_matrix_projections``) -- tiny host numpy on realisations that ``measure_xi_w`` / ``measure_xi_multipoles``
already stored; the last step of BASELINE.json config 5.  In the reference these live on ``MeasureJackknife`` (not on
``MeasureIABox``); here they are a mixin available on both ``MeasureIABox`` and a light ``MeasureJackknife``.
"""
import numpy as np

from .io import create_group_hdf5, open_file, write_dataset_hdf5

_VALID = ("w_g_plus", "multipoles_g_plus", "w_gg", "multipoles_gg")


class JackknifeCombinationMixin:
	def measure_covariance_multiple_datasets(self, corr_type, dataset_names, num_box=27, return_output=False, _handle=None):
		"""Jackknife covariance of one dataset with itself or of two datasets with each other
		(measure_jackknife.py:485-571): cov[:, i] = (n-1)/n sum_b (x_b - mean_x) (y_b[i] - mean_y[i]).
		`_handle` (not in the reference): an output file the caller keeps open, used instead of opening / closing it here."""
		if corr_type not in _VALID:
			raise ValueError("corr_type must be 'w_g_plus', 'w_gg', 'multipoles_g_plus' or 'multipoles_gg'.")
		if len(dataset_names) not in (1, 2):
			raise KeyError("Too many datasets given, choose either 1 or 2")
		f = _handle if _handle is not None else open_file(self.output_file_name, "a")
		try:
			reals = []
			for name in dataset_names:
				grp = f[f"{self.snap_group}/{corr_type}/{name}_jk{num_box}"]
				reals.append(np.array([grp[f"{name}_{b}"][:] for b in range(num_box)]))
		finally:
			if _handle is None:
				f.close()
		means = []
		for x in reals:
			m = np.zeros(self.num_bins_r)
			for b in range(num_box):
				m += x[b]
			m /= num_box
			means.append(m)
		x, mx = reals[0], means[0]
		y, my = (reals[0], means[0]) if len(reals) == 1 else (reals[1], means[1])
		cov = np.zeros((self.num_bins_r, self.num_bins_r))
		std = np.zeros(self.num_bins_r)
		for b in range(num_box):
			dx, dy = x[b] - mx, y[b] - my
			std += dx ** 2 if len(reals) == 1 else dx * dy
			for i in range(self.num_bins_r):
				cov[:, i] += dx * dy[i]
		std *= (num_box - 1) / num_box
		with np.errstate(invalid="ignore"):
			std = np.sqrt(std)
		cov *= (num_box - 1) / num_box
		if self.output_file_name is not None and not return_output:
			f = _handle if _handle is not None else open_file(self.output_file_name, "a")
			try:
				grp = create_group_hdf5(f, f"{self.snap_group}/{corr_type}")
				stem = dataset_names[0] if len(dataset_names) == 1 else dataset_names[0] + "_" + dataset_names[1]
				write_dataset_hdf5(grp, f"{stem}_jackknife_cov_{num_box}", data=cov)
				write_dataset_hdf5(grp, f"{stem}_jackknife_{num_box}", data=std)
			finally:
				if _handle is None:
					f.close()
			return None
		return cov, std

	def create_full_cov_matrix_projections(self, corr_type, dataset_names=["LOS_x", "LOS_y", "LOS_z"], num_box=27,
										   return_output=False, _handle=None):
		"""Block covariance of three projections and of each pair of projections (measure_jackknife.py:573-648).

		Kept quirk of the reference: the blocks it calls `cov_yz` / `cov_xz` are read from the datasets
		`<x>_<z>_jackknife_cov` / `<y>_<z>_jackknife_cov` respectively (:606-607), i.e. swapped; the assembled matrices
		below use them exactly as the reference does, so the stored results are identical."""
		n0, n1, n2 = dataset_names
		for pair in ((n0, n1), (n0, n2), (n1, n2)):
			self.measure_covariance_multiple_datasets(corr_type=corr_type, dataset_names=list(pair), num_box=num_box,
													  _handle=_handle)
		f = _handle if _handle is not None else open_file(self.output_file_name, "a")
		try:
			grp = f[f"{self.snap_group}/{corr_type}"]
			cov_xx = grp[f"{n0}_jackknife_cov_{num_box}"][:]
			cov_yy = grp[f"{n1}_jackknife_cov_{num_box}"][:]
			cov_zz = grp[f"{n2}_jackknife_cov_{num_box}"][:]
			cov_xy = grp[f"{n0}_{n1}_jackknife_cov_{num_box}"][:]
			cov_yz = grp[f"{n0}_{n2}_jackknife_cov_{num_box}"][:]  # sic (reference :606)
			cov_xz = grp[f"{n1}_{n2}_jackknife_cov_{num_box}"][:]  # sic (reference :607)
			cov3 = np.concatenate((np.concatenate((cov_xx, cov_xy, cov_xz), axis=1),
								   np.concatenate((cov_xy.T, cov_yy, cov_yz), axis=1),
								   np.concatenate((cov_xz.T, cov_yz.T, cov_zz), axis=1)), axis=0)
			two = lambda a, ab, b: np.concatenate((np.concatenate((a, ab), axis=1), np.concatenate((ab.T, b), axis=1)), axis=0)  # noqa: E731
			cov2xy, cov2xz, cov2yz = two(cov_xx, cov_xy, cov_yy), two(cov_xx, cov_xz, cov_zz), two(cov_yy, cov_yz, cov_zz)
			if return_output:
				return cov3, cov2xy, cov2xz, cov2yz
			write_dataset_hdf5(grp, f"{n0}_{n1}_{n2}_combined_jackknife_cov_{num_box}", data=cov3)
			write_dataset_hdf5(grp, f"{n0}_{n1}_combined_jackknife_cov_{num_box}", data=cov2xy)
			write_dataset_hdf5(grp, f"{n0}_{n2}_combined_jackknife_cov_{num_box}", data=cov2xz)
			write_dataset_hdf5(grp, f"{n1}_{n2}_combined_jackknife_cov_{num_box}", data=cov2yz)
		finally:
			if _handle is None:
				f.close()
		return None

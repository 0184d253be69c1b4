"""Run the UNMODIFIED reference (`/root/reference/src/measureia`) in this container.  TEST INFRASTRUCTURE ONLY.

`/root/reference` does not exist on the GPU box, so nothing here is used by `-m gpu` tests, `smoke()` or `bench.py`.
Uses: (1) validate the C restatement in `oracle/oracle.c` (tests/test_oracle_vs_reference.py, skipped when the
reference tree is absent); (2) generate the committed fixtures `tests/golden/*.npz` (oracle/make_golden.py).
"""
import os
import sys
import tempfile

import numpy as np

REFERENCE_SRC = "/root/reference/src"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def reference_available():
	return os.path.isdir(os.path.join(REFERENCE_SRC, "measureia"))


def _lpmn(m, n, z):
	"""`scipy.special.lpmn` was removed from scipy; re-provide it for `measure_IA_base.py:5,642`."""
	from scipy.special import assoc_legendre_p_all
	vals = assoc_legendre_p_all(n, m, z, diff_n=1)  # shape (2, n+1, 2m+1)
	p = np.array([[vals[0][j, i] for j in range(n + 1)] for i in range(m + 1)])
	dp = np.array([[vals[1][j, i] for j in range(n + 1)] for i in range(m + 1)])
	return p, dp


def load_reference():
	"""Import and return the reference package `measureia`, unmodified, with import stand-ins on sys.path."""
	if not reference_available():
		raise RuntimeError("/root/reference is not present (expected on the GPU box)")
	for p in (os.path.join(_HERE, "ref_shims"), REFERENCE_SRC, _REPO):
		if p not in sys.path:
			sys.path.insert(0, p)
	import scipy.special as sp
	if not hasattr(sp, "lpmn"):
		sp.lpmn = _lpmn
	import measureia
	return measureia


def _flatten(group, prefix=""):
	from measure_ia_b200 import h5lite
	out = {}
	for k, v in group.items():
		if isinstance(v, h5lite.Group):
			out.update(_flatten(v, prefix + k + "/"))
		else:
			out[prefix + k] = v[...]
	return out


def run_reference(data, kind, dataset_name="All", corr_type="both", num_jk=0, variant="tree", boxsize=None,
				  simulation=None, snapshot=None, separation_limits=(0.1, 20.0), num_bins_r=8, num_bins_pi=20,
				  pi_max=None, periodicity=True, masks=None, ellipticity="distortion", num_nodes=1, chunk_size=1000,
				  quiet=True):
	"""Run `MeasureIABox.measure_xi_w` (kind='w') or `.measure_xi_multipoles` (kind='multipoles') of the reference.

	variant: 'tree' (temp path given, num_nodes=1), 'brute' (temp_file_path=False), 'multiprocessing' (num_nodes>1).
	Returns {hdf5 path: ndarray} for every dataset the reference wrote.
	"""
	from measure_ia_b200 import h5lite
	measureia = load_reference()
	data = dict(data)  # the reference injects default weights into the dict it is given
	tmpdir = tempfile.mkdtemp(prefix="mia_ref_")
	out = os.path.join(tmpdir, "out.hdf5")
	devnull = open(os.devnull, "w")
	stdout = sys.stdout
	try:
		if quiet:
			sys.stdout = devnull
		nodes = num_nodes if variant == "multiprocessing" else 1
		obj = measureia.MeasureIABox(data, out, simulation, snapshot, list(separation_limits), num_bins_r,
									 num_bins_pi, pi_max, boxsize, periodicity, nodes)
		temp = False if variant == "brute" else tmpdir + "/"
		kw = dict(num_jk=num_jk, temp_file_path=temp, masks=masks, ellipticity=ellipticity, chunk_size=chunk_size)
		if kind == "w":
			obj.measure_xi_w(dataset_name, corr_type, **kw)
		else:
			obj.measure_xi_multipoles(dataset_name, corr_type, **kw)
	finally:
		sys.stdout = stdout
		devnull.close()
	f = h5lite.File(out, "r")
	res = _flatten(f)
	f.close()
	for fn in os.listdir(tmpdir):
		os.remove(os.path.join(tmpdir, fn))
	os.rmdir(tmpdir)
	return res


if __name__ == "__main__":
	import time
	sys.path.insert(0, _REPO)
	from measure_ia_b200.synthetic import uniform_box
	n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
	d = uniform_box(n, 205.0, seed=1)
	for kind in ("w", "multipoles"):
		t = time.time()
		r = run_reference(d, kind, num_jk=27, boxsize=205.0, num_bins_r=10, num_bins_pi=8)
		key = "w/xi_gg/All_DD" if kind == "w" else "multipoles/xi_gg/All_DD"
		print(kind, "datasets:", len(r), "sum DD:", r[key].sum(), f"{time.time() - t:.1f}s")

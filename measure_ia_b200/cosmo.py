"""Comoving radial distances for the light-cone path (the reference calls ``pyccl.comoving_radial_distance``,
measure_w_lightcone.py:128-130 with ``ccl.Cosmology(Omega_c=0.225, Omega_b=0.045, sigma8=0.8, h=0.7, n_s=1.0)`` as default).

``pyccl`` (pinned ~=3.2 by the reference, uv.lock: 3.2.1) is absent from this image.  When it is importable and the caller
passes one of its ``Cosmology`` objects it is used; otherwise distances come from the flat LCDM integral below (matter +
cosmological constant, no radiation / neutrinos -- CCL's default adds those: a ~1e-4 relative difference at z < 1, so
the DISTANCE conversion is not pinned against CCL; everything downstream of the distances is).
Units: Mpc (not Mpc/h), like CCL.
"""
import numpy as np

C_KM_S = 299792.458


class Cosmology:
	"""Minimal stand-in for ``pyccl.Cosmology``: keyword construction and ``cosmo["h"]`` item access."""

	def __init__(self, Omega_c=0.225, Omega_b=0.045, sigma8=0.8, h=0.7, n_s=1.0, **extra):
		self._p = dict(Omega_c=Omega_c, Omega_b=Omega_b, sigma8=sigma8, h=h, n_s=n_s, **extra)
		self._p["Omega_m"] = Omega_c + Omega_b

	def __getitem__(self, key):
		return self._p[key]


_GL = {n: np.polynomial.legendre.leggauss(n) for n in (16, 32, 64)}


def flat_lcdm_distance(omega_m, h, a):
	"""chi(a) = c / H0 int_0^z dz' / sqrt(Om (1 + z')^3 + 1 - Om), Gauss-Legendre per object on [0, z].  The integrand is
	smooth: 16 nodes are converged to rounding (2e-16 against adaptive quadrature) up to z = 1, 32 up to z = 10; beyond
	that 64 nodes (2e-15 at z = 10, degrading slowly: a light-cone catalogue does not go there).

	`a`: a numpy array, or a float64 torch tensor (any device) -- the same sequence of individually rounded IEEE operations
	either way, so the device version returns the host version's bits."""
	is_torch = type(a).__module__.startswith("torch")
	if is_torch:
		import torch
		sqrt, zeros_like = torch.sqrt, torch.zeros_like
	else:
		a = np.asarray(a, dtype=np.float64)
		sqrt, zeros_like = np.sqrt, np.zeros_like
	z = 1.0 / a - 1.0
	zmax = float(z.max()) if z.shape[0] else 0.0
	xs, ws = _GL[16 if zmax <= 1.0 else (32 if zmax <= 10.0 else 64)]
	half = 0.5 * z
	acc = zeros_like(z)
	lam = 1.0 - omega_m
	for x, w in zip(xs, ws):
		t = half * (float(x) + 1.0) + 1.0
		acc += float(w) * (1.0 / sqrt(omega_m * (t * t * t) + lam))  # (reciprocal, then product: torch evaluates scalar / tensor that way)
	return C_KM_S / (100.0 * h) * half * acc


def is_builtin_flat_lcdm(cosmology):
	"""True when `comoving_radial_distance` would use the flat-LCDM integral above (so it may as well run on the device)."""
	return isinstance(cosmology, Cosmology)


def comoving_radial_distance(cosmology, a):
	"""Same call signature as ``pyccl.comoving_radial_distance(cosmology, a)``."""
	try:  # pragma: no cover - pyccl is absent from the build image
		import pyccl
		if isinstance(cosmology, pyccl.Cosmology):
			return pyccl.comoving_radial_distance(cosmology, a)
	except Exception:  # noqa: BLE001
		pass
	if callable(cosmology):  # a user-supplied chi(a)
		return np.asarray(cosmology(np.asarray(a, dtype=np.float64)), dtype=np.float64)
	om = cosmology["Omega_c"] + cosmology["Omega_b"]
	return flat_lcdm_distance(om, cosmology["h"], a)

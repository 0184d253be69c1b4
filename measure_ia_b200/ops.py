"""PyTorch custom ops over the C ABI of libmia_b200.so (include/mia_b200.h).

``torch.ops.measure_ia_b200.paircount`` is the operator the host mirror (``box.py``) calls; it hands raw device
pointers and the current CUDA stream to ``mia_paircount`` through ctypes.  There is NO CPU implementation and no
fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch

from .build import LIB_PATH

MIA_ABI_VERSION = 3
GEOM_RPPI, GEOM_RMU = 0, 1
KERNEL_AUTO, KERNEL_GENERAL, KERNEL_TILED, KERNEL_TILED_ORDERED, KERNEL_TILED_SYM = 0, 1, 2, 3, 4
# "tiled_ordered": the tiled kernels without the symmetric auto-correlation path (every ordered pair evaluated on its own)
KERNEL_NAMES = {"auto": KERNEL_AUTO, "general": KERNEL_GENERAL, "tiled": KERNEL_TILED, "tiled_ordered": KERNEL_TILED_ORDERED}
KERNEL_REPORTED = {KERNEL_GENERAL: "general", KERNEL_TILED: "tiled", KERNEL_TILED_SYM: "tiled_sym"}


class MiaParams(ctypes.Structure):
	_fields_ = [
		("abi_version", ctypes.c_int32), ("geometry", ctypes.c_int32), ("n_r", ctypes.c_int32), ("n_2", ctypes.c_int32),
		("los", ctypes.c_int32), ("periodic", ctypes.c_int32), ("num_jk", ctypes.c_int32), ("kernel", ctypes.c_int32),
		("boxsize", ctypes.c_double), ("r_search", ctypes.c_double), ("rp2_cut", ctypes.c_double),
		("r2_thr_host", ctypes.c_void_p), ("thr2_host", ctypes.c_void_p), ("timings_host", ctypes.c_void_p),
		("variance", ctypes.c_int32),
	]


class MiaSample(ctypes.Structure):
	_fields_ = [("n", ctypes.c_int64), ("pos", ctypes.c_void_p), ("weight", ctypes.c_void_p), ("jk", ctypes.c_void_p),
				("axis", ctypes.c_void_p), ("e", ctypes.c_void_p)]


class MiaHist(ctypes.Structure):
	_fields_ = [("dd_count", ctypes.c_void_p), ("dd_w", ctypes.c_void_p), ("spd", ctypes.c_void_p),
				("scd", ctypes.c_void_p), ("dd_jk_count", ctypes.c_void_p), ("dd_jk_w", ctypes.c_void_p),
				("spd_jk", ctypes.c_void_p), ("stats", ctypes.c_void_p), ("var", ctypes.c_void_p)]


class MiaShard(ctypes.Structure):
	_fields_ = [("index", ctypes.c_int32), ("count", ctypes.c_int32)]


class MiaLcParams(ctypes.Structure):
	_fields_ = [("abi_version", ctypes.c_int32), ("geometry", ctypes.c_int32), ("n_r", ctypes.c_int32), ("n_2", ctypes.c_int32),
				("num_patches", ctypes.c_int32), ("shapes", ctypes.c_int32), ("proj_scale", ctypes.c_double),
				("rp2_cut", ctypes.c_double), ("r2_thr_host", ctypes.c_void_p), ("thr2_host", ctypes.c_void_p),
				("timings_host", ctypes.c_void_p)]


class MiaLcSample(ctypes.Structure):
	_fields_ = [("n", ctypes.c_int64), ("ra", ctypes.c_void_p), ("dec", ctypes.c_void_p), ("chi", ctypes.c_void_p),
				("cosdec", ctypes.c_void_p), ("weight", ctypes.c_void_p), ("e1", ctypes.c_void_p), ("e2", ctypes.c_void_p),
				("patch", ctypes.c_void_p)]


KERNEL_LIGHTCONE = 5
EXPORTS = ("mia_strerror", "mia_abi_version", "mia_workspace_bytes", "mia_paircount", "mia_paircount_host",
		   "mia_combine_partials_f64", "mia_lightcone_paircount", "mia_lightcone_paircount_host")

_lib = None


def load_library():
	"""dlopen libmia_b200.so; raises (never falls back) when it is missing or has the wrong ABI."""
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
							   "(there is no CPU fallback for the pair-count operator)")
		from . import build
		if build._stale():
			# built from other sources than the ones in the tree (same ABI number, different kernels): rebuild, or refuse
			import shutil
			import warnings
			if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
				warnings.warn("libmia_b200.so does not match the sources in the tree: rebuilding")
				build.build_library()
			else:
				raise RuntimeError("libmia_b200.so does not match the sources in the tree and nvcc is not available to rebuild")
		lib = ctypes.CDLL(LIB_PATH)
		lib.mia_strerror.restype = ctypes.c_char_p
		lib.mia_strerror.argtypes = [ctypes.c_int]
		lib.mia_abi_version.restype = ctypes.c_int
		lib.mia_workspace_bytes.restype = ctypes.c_size_t
		lib.mia_workspace_bytes.argtypes = [ctypes.POINTER(MiaParams), ctypes.c_int64, ctypes.c_int64]
		lib.mia_paircount.restype = ctypes.c_int
		lib.mia_paircount.argtypes = [ctypes.POINTER(MiaParams), ctypes.POINTER(MiaSample), ctypes.POINTER(MiaSample),
									  MiaShard, ctypes.POINTER(MiaHist), ctypes.c_void_p, ctypes.c_size_t,
									  ctypes.c_void_p]
		lib.mia_paircount_host.restype = ctypes.c_int
		lib.mia_paircount_host.argtypes = [ctypes.POINTER(MiaParams), ctypes.POINTER(MiaSample),
										   ctypes.POINTER(MiaSample), MiaShard, ctypes.POINTER(MiaHist), ctypes.c_int]
		lib.mia_combine_partials_f64.restype = ctypes.c_int
		lib.mia_combine_partials_f64.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p,
												 ctypes.c_void_p]
		lib.mia_lightcone_paircount.restype = ctypes.c_int
		lib.mia_lightcone_paircount.argtypes = [ctypes.POINTER(MiaLcParams), ctypes.POINTER(MiaLcSample),
												ctypes.POINTER(MiaLcSample), MiaShard, ctypes.POINTER(MiaHist), ctypes.c_void_p]
		lib.mia_lightcone_paircount_host.restype = ctypes.c_int
		lib.mia_lightcone_paircount_host.argtypes = [ctypes.POINTER(MiaLcParams), ctypes.POINTER(MiaLcSample),
													 ctypes.POINTER(MiaLcSample), MiaShard, ctypes.POINTER(MiaHist), ctypes.c_int]
		if lib.mia_abi_version() != MIA_ABI_VERSION:
			raise RuntimeError("libmia_b200.so ABI version mismatch; rebuild")
		_lib = lib
	return _lib


def check(rc):
	if rc != 0:
		raise RuntimeError(f"libmia_b200: {load_library().mia_strerror(rc).decode()} (code {rc})")


LAST_TIMINGS_MS = [0.0, 0.0, 0.0, 0.0]  # [cell-list build, pair kernel, reductions, whole call] of the last paircount call
_timing_buf = (ctypes.c_float * 4)()


def make_params(geometry, n_r, n_2, los, periodic, num_jk, kernel, boxsize, r_search, rp2_cut, r2_thr, thr2, timings=None,
				variance=False):
	"""r2_thr / thr2: CPU float64 tensors (kept alive by the caller for the duration of the call)."""
	assert r2_thr.dtype == torch.float64 and thr2.dtype == torch.float64 and not r2_thr.is_cuda and not thr2.is_cuda
	assert r2_thr.numel() == n_r + 1 and thr2.numel() == n_2 + 1
	return MiaParams(MIA_ABI_VERSION, geometry, n_r, n_2, los, 1 if periodic else 0, num_jk, kernel, boxsize, r_search,
					 rp2_cut, r2_thr.data_ptr(), thr2.data_ptr(), ctypes.addressof(timings) if timings is not None else None,
					 1 if variance else 0)


def _dev_ptr(t: Optional[torch.Tensor], dtype, shape_tail=None):
	if t is None:
		return None
	if not t.is_cuda:
		raise RuntimeError("measure_ia_b200::paircount needs CUDA tensors (no CPU fallback)")
	if t.dtype != dtype or not t.is_contiguous():
		raise RuntimeError(f"expected a contiguous {dtype} tensor, got {t.dtype} (contiguous={t.is_contiguous()})")
	if shape_tail is not None and tuple(t.shape[1:]) != tuple(shape_tail):
		raise RuntimeError(f"expected shape (N, {', '.join(map(str, shape_tail))}), got {tuple(t.shape)}" if shape_tail
						   else f"expected a 1-D tensor, got shape {tuple(t.shape)}")
	return t.data_ptr()


def _check_rows(name, t, n):
	if t is not None and t.shape[0] != n:
		raise RuntimeError(f"{name}: {t.shape[0]} rows, expected {n}")


@torch.library.custom_op("measure_ia_b200::paircount", mutates_args=())
def paircount(pos_d: torch.Tensor, weight_d: Optional[torch.Tensor], jk_d: Optional[torch.Tensor],
			  pos_s: torch.Tensor, weight_s: Optional[torch.Tensor], jk_s: Optional[torch.Tensor],
			  axis: torch.Tensor, e: torch.Tensor, r2_thr: torch.Tensor, thr2: torch.Tensor, geometry: int, los: int,
			  periodic: bool, num_jk: int, boxsize: float, r_search: float, rp2_cut: float, kernel: int,
			  shard_index: int, shard_count: int, variance: bool = False) -> List[torch.Tensor]:
	"""Binned pair sums of the position sample ``*_d`` around the shape sample ``*_s``.

	Returns [dd_count i64 (n_r,n_2), dd_w, spd, scd, dd_jk_count i64 (num_jk,n_r,n_2), dd_jk_w, spd_jk, stats u64->i64 (8),
	var (n_r,n_2) = sum (w_D w_S e+)^2 when ``variance`` else an empty tensor].
	Semantics: include/mia_b200.h; reference seam: measure_w_box_jk.py:646 / measure_m_box_jk.py:682.
	"""
	lib = load_library()
	dev = pos_d.device
	n_r, n_2 = r2_thr.numel() - 1, thr2.numel() - 1
	params = make_params(geometry, n_r, n_2, los, periodic, num_jk, kernel, boxsize, r_search, rp2_cut, r2_thr, thr2,
						 _timing_buf, variance)
	f64, i32, i64 = torch.float64, torch.int32, torch.int64
	# shapes are validated here, at the operator boundary: the library indexes pos as [n][3] and axis as [n][2]
	D = MiaSample(pos_d.shape[0], _dev_ptr(pos_d, f64, (3,)), _dev_ptr(weight_d, f64, ()), _dev_ptr(jk_d, i32, ()), None, None)
	S = MiaSample(pos_s.shape[0], _dev_ptr(pos_s, f64, (3,)), _dev_ptr(weight_s, f64, ()), _dev_ptr(jk_s, i32, ()),
				  _dev_ptr(axis, f64, (2,)), _dev_ptr(e, f64, ()))
	for nm, t, n in (("weight_d", weight_d, D.n), ("jk_d", jk_d, D.n), ("weight_s", weight_s, S.n), ("jk_s", jk_s, S.n),
					 ("axis", axis, S.n), ("e", e, S.n)):
		_check_rows(nm, t, n)
	with torch.cuda.device(dev):
		dd_count = torch.empty((n_r, n_2), dtype=i64, device=dev)
		dd_w, spd, scd = (torch.empty((n_r, n_2), dtype=f64, device=dev) for _ in range(3))
		jk_count = torch.empty((num_jk, n_r, n_2), dtype=i64, device=dev)
		jk_w, spd_jk = (torch.empty((num_jk, n_r, n_2), dtype=f64, device=dev) for _ in range(2))
		stats = torch.zeros(8, dtype=i64, device=dev)
		var = torch.empty((n_r, n_2) if variance else (0,), dtype=f64, device=dev)
		ws_bytes = lib.mia_workspace_bytes(ctypes.byref(params), D.n, S.n)
		if ws_bytes == 0:
			raise RuntimeError("libmia_b200: mia_workspace_bytes rejected the parameters")
		ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
		H = MiaHist(dd_count.data_ptr(), dd_w.data_ptr(), spd.data_ptr(), scd.data_ptr(),
					jk_count.data_ptr() if num_jk else None, jk_w.data_ptr() if num_jk else None,
					spd_jk.data_ptr() if num_jk else None, stats.data_ptr(), var.data_ptr() if variance else None)
		stream = torch.cuda.current_stream(dev).cuda_stream
		rc = lib.mia_paircount(ctypes.byref(params), ctypes.byref(D), ctypes.byref(S), MiaShard(shard_index, shard_count),
							   ctypes.byref(H), ws.data_ptr(), ws_bytes, stream)
	check(rc)
	LAST_TIMINGS_MS[:] = list(_timing_buf)
	return [dd_count, dd_w, spd, scd, jk_count, jk_w, spd_jk, stats, var]


@paircount.register_fake
def _(pos_d, weight_d, jk_d, pos_s, weight_s, jk_s, axis, e, r2_thr, thr2, geometry, los, periodic, num_jk, boxsize,
	  r_search, rp2_cut, kernel, shard_index, shard_count, variance=False):
	n_r, n_2 = r2_thr.numel() - 1, thr2.numel() - 1
	f = lambda *s, dt=torch.float64: torch.empty(s, dtype=dt, device=pos_d.device)  # noqa: E731
	return [f(n_r, n_2, dt=torch.int64), f(n_r, n_2), f(n_r, n_2), f(n_r, n_2), f(num_jk, n_r, n_2, dt=torch.int64),
			f(num_jk, n_r, n_2), f(num_jk, n_r, n_2), f(8, dt=torch.int64), f(n_r, n_2) if variance else f(0)]


def paircount_host(params: MiaParams, D: MiaSample, S: MiaSample, H: MiaHist, shard=(0, 1), device=0):
	"""``mia_paircount_host``: host pointers in, host pointers out (what a non-torch binding would call)."""
	check(load_library().mia_paircount_host(ctypes.byref(params), ctypes.byref(D), ctypes.byref(S),
											MiaShard(shard[0], shard[1]), ctypes.byref(H), device))


def combine_partials(parts: torch.Tensor) -> torch.Tensor:
	"""Fixed-order sum over dim 0 of a [n_parts, ...] float64 CUDA tensor (measure_w_box_jk.py:775-780)."""
	if not parts.is_cuda or parts.dtype != torch.float64:
		raise RuntimeError("combine_partials needs a float64 CUDA tensor")
	parts = parts.contiguous()
	out = torch.empty(parts.shape[1:], dtype=torch.float64, device=parts.device)
	with torch.cuda.device(parts.device):
		check(load_library().mia_combine_partials_f64(parts.data_ptr(), parts.shape[0], out.numel(), out.data_ptr(),
													  torch.cuda.current_stream(parts.device).cuda_stream))
	return out


# ---- light-cone brute pair loops (include/mia_b200.h, mia_lightcone_paircount) -------------------------------------------
LAST_LC_TIMINGS_MS = [0.0, 0.0]  # [pair kernel, whole call] of the last lightcone_paircount call
_lc_timing_buf = (ctypes.c_float * 2)()


def lightcone_paircount(position: dict, shape: dict, r2_thr: torch.Tensor, thr2: torch.Tensor, geometry: int, shapes: bool,
						num_patches: int = 0, proj_scale: float = 1.0, rp2_cut: float = 0.0, shard_index: int = 0,
						shard_count: int = 1) -> List[torch.Tensor]:
	"""Pair sums of a light-cone position sample around a shape sample (reference loops: measure_w_lightcone.py:137-183,
	:307-334, measure_m_lightcone.py:142-191, :300-330).

	position: dict of 1-D float64 CUDA tensors ra, dec, chi, cosdec[, weight][, patch (int32)], SORTED by chi;
	shape: ra, dec, chi[, weight][, e1, e2 (= e cos 2phi_axis, e sin 2phi_axis)][, patch].  No CPU implementation.
	Returns [dd_count i64 (n_r,n_2), dd_w, spd, scd, touch_count i64 (num_patches,n_r,n_2), touch_w, touch_spd, stats i64 (8)].
	"""
	lib = load_library()
	dev = position["chi"].device
	n_r, n_2 = r2_thr.numel() - 1, thr2.numel() - 1
	assert r2_thr.dtype == torch.float64 and thr2.dtype == torch.float64 and not r2_thr.is_cuda and not thr2.is_cuda
	params = MiaLcParams(MIA_ABI_VERSION, geometry, n_r, n_2, num_patches, 1 if shapes else 0, float(proj_scale), float(rp2_cut),
						 r2_thr.data_ptr(), thr2.data_ptr(), ctypes.addressof(_lc_timing_buf))
	f64, i32, i64 = torch.float64, torch.int32, torch.int64

	def sample(d, is_shape):
		n = d["chi"].shape[0]
		for k, t in d.items():
			if t is not None:
				_check_rows(k, t, n)
		return MiaLcSample(n, _dev_ptr(d["ra"], f64, ()), _dev_ptr(d["dec"], f64, ()), _dev_ptr(d["chi"], f64, ()),
						   None if is_shape else _dev_ptr(d["cosdec"], f64, ()), _dev_ptr(d.get("weight"), f64, ()),
						   _dev_ptr(d.get("e1"), f64, ()) if is_shape else None, _dev_ptr(d.get("e2"), f64, ()) if is_shape else None,
						   _dev_ptr(d.get("patch"), i32, ()))

	D, S = sample(position, False), sample(shape, True)
	with torch.cuda.device(dev):
		dd_count = torch.empty((n_r, n_2), dtype=i64, device=dev)
		dd_w, spd, scd = (torch.empty((n_r, n_2), dtype=f64, device=dev) for _ in range(3))
		t_count = torch.empty((num_patches, n_r, n_2), dtype=i64, device=dev)
		t_w, t_spd = (torch.empty((num_patches, n_r, n_2), dtype=f64, device=dev) for _ in range(2))
		stats = torch.zeros(8, dtype=i64, device=dev)
		H = MiaHist(dd_count.data_ptr(), dd_w.data_ptr(), spd.data_ptr(), scd.data_ptr(),
					t_count.data_ptr() if num_patches else None, t_w.data_ptr() if num_patches else None,
					t_spd.data_ptr() if num_patches else None, stats.data_ptr(), None)
		stream = torch.cuda.current_stream(dev).cuda_stream
		rc = lib.mia_lightcone_paircount(ctypes.byref(params), ctypes.byref(D), ctypes.byref(S),
										 MiaShard(shard_index, shard_count), ctypes.byref(H), stream)
	check(rc)
	LAST_LC_TIMINGS_MS[:] = list(_lc_timing_buf)
	return [dd_count, dd_w, spd, scd, t_count, t_w, t_spd, stats]


def lightcone_paircount_host(params: MiaLcParams, D: MiaLcSample, S: MiaLcSample, H: MiaHist, shard=(0, 1), device=0):
	"""``mia_lightcone_paircount_host``: host pointers in, host pointers out."""
	check(load_library().mia_lightcone_paircount_host(ctypes.byref(params), ctypes.byref(D), ctypes.byref(S),
													  MiaShard(shard[0], shard[1]), ctypes.byref(H), device))

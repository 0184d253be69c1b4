"""Build / locate libmia_b200.so (in-tree: measure_ia_b200/lib/, so the built file travels with the repository)."""
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIA_LIB_PATH", os.path.join(_HERE, "lib", "libmia_b200.so"))
PEAKS_PATH = os.path.join(_HERE, "lib", "libmia_peaks.so")
HASH_PATH = os.path.join(_HERE, "lib", "libmia_b200.srchash")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "mia_b200.h")


def source_hash():
	"""sha256 over the CUDA sources, the Makefile and the C header (content, not mtimes: the repository is copied to the
	GPU box, which scrambles timestamps)."""
	h = hashlib.sha256()
	files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) or f == "Makefile")
	for path in files + [HEADER]:
		h.update(os.path.basename(path).encode())
		with open(path, "rb") as fh:
			h.update(fh.read())
	return h.hexdigest()


def _stale():
	"""True when the library is missing or was built from other sources than the ones in the tree."""
	if "MIA_LIB_PATH" in os.environ:  # an explicitly chosen library (tuning variants) is taken as is
		return not os.path.exists(LIB_PATH)
	if not os.path.exists(LIB_PATH) or not os.path.exists(PEAKS_PATH) or not os.path.exists(HASH_PATH):
		return True
	with open(HASH_PATH) as fh:
		return fh.read().strip() != source_hash()


def build_library(force=False, verbose=False):
	"""nvcc -gencode arch=compute_100a,code=sm_100a ... -> measure_ia_b200/lib/libmia_b200.so"""
	if force or _stale():
		out = None if verbose else subprocess.DEVNULL
		subprocess.check_call(["make", "-C", CSRC, "all", "-B"], stdout=out)
		with open(HASH_PATH, "w") as fh:
			fh.write(source_hash() + "\n")
	return LIB_PATH

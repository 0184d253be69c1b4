"""Build / locate libmia_b200.so (in-tree: measure_ia_b200/lib/, so the built file travels with the repository)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIA_LIB_PATH", os.path.join(_HERE, "lib", "libmia_b200.so"))
PEAKS_PATH = os.path.join(_HERE, "lib", "libmia_peaks.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "mia_b200.h")


def _stale():
	if not os.path.exists(LIB_PATH) or not os.path.exists(PEAKS_PATH):
		return True
	t = os.path.getmtime(LIB_PATH)
	srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + [HEADER]
	return any(os.path.getmtime(s) > t for s in srcs)


def build_library(force=False, verbose=False):
	"""nvcc -gencode arch=compute_100a,code=sm_100a ... -> measure_ia_b200/lib/libmia_b200.so"""
	if force or _stale():
		out = None if verbose else subprocess.DEVNULL
		subprocess.check_call(["make", "-C", CSRC, "all"] + (["-B"] if force else []), stdout=out)
	return LIB_PATH

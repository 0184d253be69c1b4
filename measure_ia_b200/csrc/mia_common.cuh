// mia_common.cuh -- shared device-side definitions for the B200 pair-counting kernels (sm_100a).
//
// Exactness: the reference (src/measureia/measure_w_box_jk.py:401-434, measure_m_box_jk.py:418-460) is plain IEEE
// double arithmetic evaluated by numpy, never fused.  Everything that decides WHICH bin a pair falls in is therefore
// written with the __d*_rn intrinsics (never contracted into FMA by nvcc) in the reference's order, and compared
// against threshold tables calibrated on the host with numpy (measure_ia_b200/calib.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mia_b200.h"

#define MIA_CUDA_CHECK(expr)                         \
	do {                                             \
		cudaError_t _e = (expr);                     \
		if (_e != cudaSuccess) return (int)_e;       \
	} while (0)

namespace mia {

// Candidate ("position sample") galaxy in canonical axis order: u, v = the two projected axes in ascending column
// order (not_LOS, measure_w_box_jk.py:356), l = line of sight.  32 bytes: one cell is one contiguous, 16B-aligned
// byte range, which is what lets a tile be fetched with a single bulk (TMA) copy.
struct __align__(32) Cand {
	double u, v, l, w;
};

// Primary ("shape sample") galaxy.  64 bytes.
struct __align__(32) Prim {
	double u, v, l, w;
	double a0, a1, e;  // normalised axis direction (in u, v order) and ellipticity size
	int32_t jk;        // jackknife label
	int32_t orig;      // index in the caller's arrays
};

// Kernel parameter block (passed by value).
struct DevParams {
	int geom, n_r, n_2, los, periodic, num_jk;
	double L, halfL, rp2_cut;
	double r2_thr[MIA_MAX_BINS + 1];
	double thr2[MIA_MAX_BINS + 1];
	// cell grid over (u, v, l)
	int ncu, ncv, ncl;
	double inv_cu, inv_cv, inv_cl;
	int ku, kv, kl;  // neighbour reach in cells; a reach >= nc means "all cells of that axis"
};

struct Grid {
	const Cand *cand;          // [nD] sorted by (cell, jk label)
	const int32_t *cand_jk;    // [nD]
	const int64_t *cell_start; // [ncell + 1]
	int64_t n_cand;
	int64_t n_cell;
};

// Global accumulators (device).  JK rows are kept as two families so that the total is a by-product:
//   A[k] = pairs whose SHAPE galaxy is in region k;   B[k] = pairs whose POSITION galaxy is in region k != shape's.
//   total = sum_k A[k];   jk[k] = A[k] + B[k]   (measure_w_box_jk.py:442-461)
// Without jackknife a single row A[0] is used.
struct Accum {
	unsigned long long *cnt;  // [rows][nb]
	double *ddw, *sp, *sc;    // [rows][nb]
	double *var;              // [nb] per accumulator copy (no jackknife rows): sum (w_D w_S e+)^2, or NULL
	unsigned long long *stats;
	int rows;                 // 2 * max(num_jk, 1)
};

__device__ __forceinline__ int cell_index(double x, double inv, int nc) {
	int c = (int)floor(x * inv);
	return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

// Separation along one axis, exactly as the reference forms it: shape minus position, then the two conditional
// shifts in sequence (measure_w_box_jk.py:401-404).
__device__ __forceinline__ double sep_axis(double s, double c, const DevParams &P) {
	double d = __dsub_rn(s, c);
	if (P.periodic) {
		if (d > P.halfL) d = __dsub_rn(d, P.L);
		if (d < -P.halfL) d = __dadd_rn(d, P.L);
	}
	return d;
}

// Sum of squares over the three ORIGINAL columns 0, 1, 2 in numpy's order ((x0^2 + x1^2) + x2^2),
// measure_m_box_jk.py:428, given the canonical components.
__device__ __forceinline__ double r3_squared(double du, double dv, double dl, int los) {
	double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dl, dl);
	if (los == 2) return __dadd_rn(__dadd_rn(uu, vv), ll);  // columns (u, v, l)
	if (los == 1) return __dadd_rn(__dadd_rn(uu, ll), vv);  // columns (u, l, v)
	return __dadd_rn(__dadd_rn(ll, uu), vv);                // columns (l, u, v)
}

struct PairResult {
	int rbin, bin2;
	double gp, gc;  // e_+ / e and e_x / e for this pair (0 under the NaN rule)
	bool nan_rule;
};

// Number of interior thresholds passed: the reference's bin index (see mia_b200.h).
__device__ __forceinline__ int count_thresholds(double x, const double *thr, int n) {
	int b = 0;
	for (int k = 1; k < n; k++) b += (x >= thr[k]) ? 1 : 0;
	return b;
}

// cos(2 acos c) and sin(2 acos c) for c in [-1, 1] (measure_IA_base.py:226 with phi = arccos(c), phi in [0, pi]):
//   cos 2phi = 2c^2 - 1,  sin 2phi = 2 c sqrt(1 - c^2)   (sin phi >= 0).
__device__ __forceinline__ void shape_projection(double c, double &gp, double &gc) {
	gp = fma(2.0 * c, c, -1.0);
	gc = 2.0 * c * sqrt(fma(-c, c, 1.0));
}

// Reference-exact evaluation of one (shape, position) pair.  Returns false when the pair is not binned.
template <int GEOM>
__device__ __forceinline__ bool eval_pair_exact(const DevParams &P, double su, double sv, double sl, double a0,
												double a1, double cu, double cv, double cl, PairResult &out) {
	const double du = sep_axis(su, cu, P), dv = sep_axis(sv, cv, P), dl = sep_axis(sl, cl, P);
	const double rp2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));  // measure_w_box_jk.py:407
	double s;
	if (GEOM == MIA_GEOM_RPPI) {
		s = rp2;
		if (!(s >= P.r2_thr[0] && s < P.r2_thr[P.n_r])) return false;
		if (!(dl >= P.thr2[0] && dl < P.thr2[P.n_2])) return false;  // :420-421
		out.bin2 = count_thresholds(dl, P.thr2, P.n_2);
	} else {
		if (!(rp2 > P.rp2_cut)) return false;  // measure_m_box_jk.py:444
		s = r3_squared(du, dv, dl, P.los);
		if (!(s >= P.r2_thr[0] && s < P.r2_thr[P.n_r])) return false;
		const double mu = __ddiv_rn(dl, __dsqrt_rn(s));  // :431
		out.bin2 = count_thresholds(mu, P.thr2, P.n_2);
	}
	out.rbin = count_thresholds(s, P.r2_thr, P.n_r);
	const double rp = __dsqrt_rn(rp2);
	const double d0 = __ddiv_rn(du, rp), d1 = __ddiv_rn(dv, rp);         // :409
	const double c = __dadd_rn(__dmul_rn(d0, a0), __dmul_rn(d1, a1));    // :412-413
	if (fabs(c) <= 1.0) {
		shape_projection(c, out.gp, out.gc);
		out.nan_rule = false;
	} else {  // arccos -> NaN -> e_+ = e_x = 0, the pair still counts in DD (:416-417)
		out.gp = 0.0;
		out.gc = 0.0;
		out.nan_rule = true;
	}
	return true;
}

}  // namespace mia

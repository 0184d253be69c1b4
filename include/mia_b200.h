/*
 * mia_b200.h -- C ABI of the B200 pair-counting library (libmia_b200.so).
 *
 * This is the drop-in boundary for the reference's periodic-box pair loop.  The reference has no FFI of its own (it is
 * pure Python); the seam this ABI replaces is the worker function of its data-parallel variants, which returns the
 * five accumulators (Splus_D, Scross_D, DD, DD_jk, Splus_D_jk):
 *     src/measureia/measure_w_box_jk.py:543-646   _measure_xi_rp_pi_box_jk_batch      -> mia_paircount, MIA_GEOM_RPPI
 *     src/measureia/measure_m_box_jk.py:570-682   _measure_xi_r_mur_box_jk_batch      -> mia_paircount, MIA_GEOM_RMU
 *     src/measureia/measure_w_box.py:412-491 / measure_m_box.py:629-717 (no jackknife) -> same, num_jk = 0
 * and the parent-side reduction of worker results (measure_w_box_jk.py:775-780) -> mia_combine_partials.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every array pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library allocates nothing persistent;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are re-entrant;
 *   - return value: 0 = ok, < 0 = MIA_ERR_* (bad argument / capacity), > 0 = a cudaError_t; never throws.
 *
 * Exactness contract (DESIGN.md "exact thresholds"): separations are formed with the reference's IEEE operation
 * sequence (no FMA contraction) and binned by comparing against threshold tables calibrated on the host with the
 * reference's own numpy expressions, so dd_count is bit-identical to the reference's DD for unit weights.
 */
#ifndef MIA_B200_H
#define MIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIA_ABI_VERSION 3
#define MIA_MAX_BINS 64 /* per axis */

enum { MIA_GEOM_RPPI = 0, MIA_GEOM_RMU = 1 };
/* kernel selection.  AUTO / TILED pick, for an auto-correlation (position and shape sample alias: same pos / weight / jk
 * pointers), the symmetric kernel that visits every unordered pair once; TILED_ORDERED never does.  TILED_SYM is only
 * REPORTED (stats[4]), not requested. */
enum { MIA_KERNEL_AUTO = 0, MIA_KERNEL_GENERAL = 1, MIA_KERNEL_TILED = 2, MIA_KERNEL_TILED_ORDERED = 3, MIA_KERNEL_TILED_SYM = 4 };
enum {
	MIA_OK = 0,
	MIA_ERR_ARG = -1,       /* NULL pointer, negative size, bins out of range ... */
	MIA_ERR_WORKSPACE = -2, /* workspace too small: call mia_workspace_bytes */
	MIA_ERR_RANGE = -3,     /* a coordinate is outside [0, boxsize) (the reference's KDTree raises here too) */
	MIA_ERR_WINDOW = -4,    /* internal consistency check of the tiled kernel failed (a pair fell outside its window) */
	MIA_ERR_UNSUPPORTED = -5,
	MIA_ERR_UNSORTED = -6   /* light-cone entry: a sample is not sorted by comoving distance (or holds a NaN distance) */
};

/* Binning and geometry.  Threshold tables are HOST pointers, copied during the call. */
typedef struct mia_params {
	int32_t abi_version; /* MIA_ABI_VERSION */
	int32_t geometry;    /* MIA_GEOM_RPPI: (r_p, Pi) grid; MIA_GEOM_RMU: (r, mu_r) grid */
	int32_t n_r;         /* 1..MIA_MAX_BINS */
	int32_t n_2;         /* 1..MIA_MAX_BINS: Pi bins or mu_r bins */
	int32_t los;         /* line-of-sight column 0..2 (data["LOS"], measure_w_box_jk.py:355) */
	int32_t periodic;    /* the reference's `periodicity` flag: wrap separations by +-boxsize (measure_w_box_jk.py:402-404) */
	int32_t num_jk;      /* number of jackknife regions, 0 = none (labels come with the samples) */
	int32_t kernel;      /* MIA_KERNEL_* */
	double boxsize;
	double r_search;     /* >= largest separation that can be binned (r_bins[-1]); sizes the cell grid */
	double rp2_cut;      /* MIA_GEOM_RMU: a pair needs r_p^2 > rp2_cut (measure_m_box_jk.py:444); else ignored */
	/* r2_thr_host[b], b = 0..n_r : a pair with squared separation s (r_p^2 or r^2, summed in the reference's order)
	 * is in range iff r2_thr[0] <= s < r2_thr[n_r] and then lies in r-bin  #{1 <= b < n_r : s >= r2_thr[b]}. */
	const double *r2_thr_host;
	/* thr2_host[b], b = 0..n_2 : same for the second axis (Pi, or mu_r = Pi / r); use -inf / +inf for "no limit". */
	const double *thr2_host;
	/* optional HOST pointer to 4 floats, filled when the call returns: milliseconds (CUDA events on `stream`) spent in
	 * [0] the cell-list build (keys, radix sort, gather, offsets, task table), [1] the pair kernel, [2] the fixed-order
	 * reductions, [3] the whole call.  NULL = do not time. */
	float *timings_host;
	/* 1: also accumulate mia_hist.var = sum (w_D w_S e_+)^2 per bin, the `variance` of the reference's brute variants
	 * (measure_w_box_jk.py:196,242; measure_m_box_jk.py:207,255).  Such calls use the ordered kernels. */
	int32_t variance;
} mia_params;

/* One catalogue.  pos is row-major [n][3] in the caller's column order.  axis / e are only read for the shape sample:
 * axis = normalised projected axis direction [n][2] (measure_w_box_jk.py:326-327), e = ellipticity size (:357-362).
 * jk = jackknife region label per galaxy (measure_IA_base.py:404-452), may be NULL when num_jk == 0. */
typedef struct mia_sample {
	int64_t n;
	const double *pos;
	const double *weight;
	const int32_t *jk;
	const double *axis;
	const double *e;
} mia_sample;

/* Outputs, all caller-allocated; the call overwrites them.  Grids are row-major [n_r][n_2]; *_jk are [num_jk][n_r][n_2]
 * and hold, for region k, the sum over pairs with the shape OR the position galaxy in region k
 * (measure_w_box_jk.py:442-461), so leave-one-out = total - jk[k].  spd / scd / spd_jk do NOT carry 1/(2R).
 * Any of the *_jk pointers may be NULL when num_jk == 0. */
typedef struct mia_hist {
	int64_t *dd_count;    /* binned ordered (position, shape) pairs: the bit-exact quantity */
	double *dd_w;         /* sum w_D w_S                         -> DD      */
	double *spd;          /* sum w_D w_S e_+                     -> S+D * 2R */
	double *scd;          /* sum w_D w_S e_x                     -> SxD * 2R */
	int64_t *dd_jk_count;
	double *dd_jk_w;      /* -> DD_jk      */
	double *spd_jk;       /* -> Splus_D_jk */
	uint64_t *stats;      /* [8]: 0 separations computed (the symmetric kernel computes one per UNORDERED pair), 1 ordered pairs
	                                binned, 2 |c|>1 pairs (NaN rule), 3 window errors,
	                                4 kernel used (MIA_KERNEL_*), 5 cells, 6 warp tasks, 7 kernels launched by the library
	                                (its own kernels; the CUB radix-sort / scan launches are not counted) */
	double *var;          /* [n_r][n_2] sum (w_D w_S e_+)^2 -> variance * (2R)^2; written iff params->variance (may be NULL otherwise) */
} mia_hist;

/* Shard of the shape sample handled by this call (multi-GPU: rank r of w takes the r-th of w work-balanced slices of
 * the cell-sorted shape sample; every rank holds the full position sample).  {0, 1} = everything. */
typedef struct mia_shard {
	int32_t index;
	int32_t count;
} mia_shard;

const char *mia_strerror(int code);
int mia_abi_version(void);

/* Bytes of device workspace mia_paircount needs for these sizes (0 on bad arguments). */
size_t mia_workspace_bytes(const mia_params *params, int64_t n_position, int64_t n_shape);

/* The pair loop.  position = "D" sample, shape = "S" sample (may alias for an auto-correlation). */
int mia_paircount(const mia_params *params, const mia_sample *position, const mia_sample *shape, mia_shard shard,
				  const mia_hist *out, void *workspace, size_t workspace_bytes, void *stream);

/* Same call with HOST buffers: allocates device memory, copies inputs in and results out, synchronises.
 * This is what a ctypes / cffi binding without a device-array library would call. */
int mia_paircount_host(const mia_params *params, const mia_sample *position_host, const mia_sample *shape_host,
					   mia_shard shard, const mia_hist *out_host, int device);

/* Fixed-order sum of `n_parts` partial results laid out back to back (the parent-side reduction of the reference,
 * measure_w_box_jk.py:775-780; used after an all-gather so the fp64 sums do not depend on arrival order).
 * parts is [n_parts][n_values] row-major, out is [n_values]: out[i] = ((parts[0][i] + parts[1][i]) + ...). */
int mia_combine_partials_f64(const double *parts, int32_t n_parts, int64_t n_values, double *out, void *stream);

/* ---- light-cone brute pair loops (additive to ABI v3) ---------------------------------------------------
 *
 * Replaces the O(N_p N_s) loops over position galaxies of the reference's light-cone estimators:
 *     src/measureia/measure_w_lightcone.py:137-183   _measure_xi_rp_pi_lightcone_brute     -> MIA_GEOM_RPPI, shapes = 1
 *     src/measureia/measure_w_lightcone.py:307-334   _count_pairs_xi_rp_pi_lightcone_brute -> MIA_GEOM_RPPI, shapes = 0
 *     src/measureia/measure_m_lightcone.py:142-191   _measure_xi_r_mur_lightcone_brute     -> MIA_GEOM_RMU,  shapes = 1
 *     src/measureia/measure_m_lightcone.py:300-330   _count_pairs_xi_r_mur_lightcone_brute -> MIA_GEOM_RMU,  shapes = 0
 * and, through num_patches, the leave-one-patch-out re-runs of measure_jackknife.py:116-170 in the same pass.
 * The flat-sky separation of a pair is formed at the POSITION galaxy's distance (see mia_lightcone.cuh for the operation
 * sequence); distances (pyccl in the reference) and the per-galaxy trigonometry are the caller's, so that they are the
 * caller's numpy bit for bit. */
typedef struct mia_lc_params {
	int32_t abi_version; /* MIA_ABI_VERSION */
	int32_t geometry;    /* MIA_GEOM_RPPI / MIA_GEOM_RMU */
	int32_t n_r;
	int32_t n_2;
	int32_t num_patches; /* jackknife patches (labels 0..num_patches-1 come with the samples), 0 = none */
	int32_t shapes;      /* 1: also sum w w e_+ and w w e_x (the `_measure_xi_*` loops); 0: pair counts only (`_count_pairs_*`) */
	double proj_scale;   /* factor applied to the PROJECTED separation (dx, dy) only: h with over_h, else 1
	                        (measure_w_lightcone.py:145-146; the distances themselves arrive already scaled, :131-133) */
	double rp2_cut;      /* MIA_GEOM_RMU: a pair needs r_p^2 > rp2_cut (measure_m_lightcone.py:172) */
	const double *r2_thr_host; /* as in mia_params */
	const double *thr2_host;
	float *timings_host; /* optional HOST pointer to 2 floats: [0] pair kernel ms, [1] whole call ms */
} mia_lc_params;

/* One light-cone catalogue, SORTED BY chi ASCENDING (MIA_ERR_UNSORTED otherwise).  All arrays have n entries. */
typedef struct mia_lc_sample {
	int64_t n;
	const double *ra;     /* degrees */
	const double *dec;    /* degrees */
	const double *chi;    /* comoving radial distance of the redshift (times h with over_h) */
	const double *cosdec; /* position sample: cos(dec / 180 * pi) as the caller computes it; NULL for the shape sample */
	const double *weight; /* NULL = unit weights */
	const double *e1;     /* shape sample with shapes = 1: e cos(2 phi_axis) */
	const double *e2;     /*                               e sin(2 phi_axis) */
	const int32_t *patch; /* jackknife patch label, NULL when num_patches == 0 */
} mia_lc_sample;

/* Results go to a mia_hist: dd_count / dd_w / spd / scd as for the box; dd_jk_count / dd_jk_w / spd_jk hold, per patch k,
 * the sums over pairs with the position OR the shape galaxy in patch k (realisation k of the reference = total - [k]);
 * stats[0] = separations computed, [1] = pairs binned, [4] = MIA_KERNEL_LIGHTCONE, [7] = kernels launched.  No workspace.
 * `shard`: slice of the (sorted) position sample this call handles. */
#define MIA_KERNEL_LIGHTCONE 5
int mia_lightcone_paircount(const mia_lc_params *params, const mia_lc_sample *position, const mia_lc_sample *shape,
							mia_shard shard, const mia_hist *out, void *stream);

/* Same call with HOST buffers (allocates, copies, synchronises). */
int mia_lightcone_paircount_host(const mia_lc_params *params, const mia_lc_sample *position_host,
								 const mia_lc_sample *shape_host, mia_shard shard, const mia_hist *out_host, int device);

#ifdef __cplusplus
}
#endif
#endif /* MIA_B200_H */

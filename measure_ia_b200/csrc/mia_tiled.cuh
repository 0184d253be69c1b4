// mia_tiled.cuh -- the TILED (r_p, Pi) pair kernel for sm_100a (cell-by-cell variant) and everything the tiled kernels
// share: planning, workspace, task table, mbarrier / bulk-copy helpers, the pair loop, the flush.  The row-streaming
// (r_p, Pi) variant lives in mia_tiled_rppi2.cuh, the (r, mu_r) kernel in mia_tiled_rmu.cuh; plan_tiled() picks.
//
// Replaces the hot loop of the reference, src/measureia/measure_w_box_jk.py:387-461 (and measure_w_box.py:313-365).
//
// Design ("windowed private histograms"):
//   * one thread owns one shape galaxy (registers); one WARP owns 32 consecutive cell-sorted shape galaxies of one
//     grid COLUMN (a cell of the two projected axes spanning the whole line of sight) and is the unit of scheduling:
//     warps never synchronise with each other after start-up;
//   * position galaxies are streamed cell by cell into shared memory by 1-D bulk (TMA) copies -- a cell is one
//     contiguous 32-byte-per-galaxy range of the sorted catalogue -- double buffered behind mbarriers, one stream per
//     warp; all 32 lanes read the same candidate (shared-memory broadcast);
//   * the grid is cut along the line of sight into slabs thinner than one Pi bin.  For one shape galaxy, all candidates
//     of a slab fall into at most TWO Pi bins; every staged cell carries ONE jackknife label; and r_p bins are visited
//     in WINDOWS of W_R bins (the top window holds ~96% of all pairs for log-spaced bins, lower windows only re-visit
//     the few nearest cells).  So between two flushes a thread touches 2 * W_R histogram slots: they are thread-private
//     shared-memory words updated with plain loads and stores -- no atomics anywhere in the pair loop, and small enough
//     (120 B per thread) that occupancy is limited by registers, not by shared memory;
//   * at a flush (window / slab / candidate-label change) the warp reduces its slots with shuffles in a fixed order
//     and adds them to its own accumulator copy in HBM (rows A[jk_shape], B[jk_position]); a final kernel sums the
//     copies in a fixed order.  Warp tasks are assigned to warps by a prefix sum of estimated work.  => results are
//     identical from run to run, which the reference's exact-scaling test (tests/test_weights.py:34-35) needs.
//
// Exactness: separations are formed with the reference's operation sequence (__d*_rn, no contraction); range and bin
// decisions are comparisons against the host-calibrated thresholds; the only approximate arithmetic is the VALUE of
// e+ / ex (reciprocal by one cubic iteration, FMA), accurate to ~1e-16, far inside the 1e-10 contract.  Pairs whose
// |cos| is within 1e-11 of 1 are re-evaluated with the reference's sequence to apply its NaN rule
// (measure_w_box_jk.py:416-417).
#pragma once
#include <cub/cub.cuh>
#include "mia_common.cuh"
#include "mia_grid.cuh"

namespace mia {

constexpr int TP = 128;          // threads per CTA
constexpr int TW = TP / 32;      // warps per CTA (independent workers)
// Tuning constants (measured on B200, cfg2: profiles/r01_tuning.md)
#ifndef MIA_CH
#define MIA_CH 96
#endif
#ifndef MIA_W_R
#define MIA_W_R 5
#endif
#ifndef MIA_UNROLL
#define MIA_UNROLL 2
#endif
#define MIA_PRAGMA(x) _Pragma(#x)
#define MIA_UNROLL_PRAGMA(n) MIA_PRAGMA(unroll n)
#ifndef MIA_MIN_CTAS
#define MIA_MIN_CTAS 3
#endif
#ifndef MIA_TASKS_PER_WARP
#define MIA_TASKS_PER_WARP 32.0
#endif
#ifndef MIA_SWP
#define MIA_SWP 0
#endif
#ifndef MIA_RPPI_V2
#define MIA_RPPI_V2 1
#endif
#ifndef MIA_RPPI_ALIGN
#define MIA_RPPI_ALIGN 1
#endif
constexpr int CH = MIA_CH;       // candidates per staged chunk
constexpr int STAGES = 2;        // per-warp double buffering
constexpr int MAX_NEIGH = 128;   // neighbour columns per task
constexpr int LUT_SIZE = 256;
constexpr int W_R = MIA_W_R;     // r bins per accumulation window
constexpr int NSLOT = 2 * W_R;   // private histogram slots per thread
constexpr int MAX_SPLIT = 8;     // a warp task may be cut into up to 8 line-of-sight parts when tasks are scarce
constexpr int SLOTS_PER_SM = 6;  // CTAs launched per SM; fixed, so that results do not depend on occupancy
#ifndef MIA_SLOT_MULT
#define MIA_SLOT_MULT 0  // 0: by catalogue size
#endif

struct LutEntry {
	double thr;  // threshold inside this entry's range of s (or +inf)
	int base;    // r-bin of the smallest s of the entry
	int pad;
};

struct CellInfo {
	double umin, umax, vmin, vmax;
	int label;  // jackknife label of the first candidate
	int nlab;   // number of label runs in the cell (1 = uniform)
};

// Bounding box of a whole column in the projected axes ((r, mu_r) kernel)
struct ColInfo {
	double umin, umax, vmin, vmax;
};

// Candidate of the symmetric (auto-correlation) kernels: position (+ weight) as in Cand, plus what the REVERSE pair needs: its
// normalised axis direction (u, v order) and w * e.  48 bytes with unit weights, 64 with weights; both are multiples of 16
// (bulk copies).  u, v, l lead the record like in Cand.
struct __align__(16) CandSU {
	double u, v, l, we;
	double a0, a1;
};
struct __align__(16) CandSW {
	double u, v, l, we;
	double a0, a1;
	double w, pad;
};

struct TiledConfig {
	int geom;        // MIA_GEOM_*
	int w_r;         // (r, mu_r): r bins per accumulation window
	int ratio;       // (r, mu_r): a shape column is ratio x ratio candidate columns wide
	int hsplit;      // (r, mu_r): a warp works on 32 / hsplit shape galaxies, hsplit candidates at a time
	int n_lr;        // (r, mu_r): line-of-sight regions (= jackknife sub-boxes per side when the slabs are aligned with them)
	int v2;          // (r_p, Pi): row-streaming kernel (mia_tiled_rppi2.cuh); n_lr then counts the regions along v
	int sym_ok;      // (r_p, Pi), row-streaming: the grid admits the symmetric auto-correlation kernel (mia_tiled_rppi2s.cuh)
	int sig;         // also accumulate sum (w_D w_S e+)^2 per bin (mia_params.variance)
	int sym;         // ... and this call uses it (position sample == shape sample): bytes per candidate record (CandSU / CandSW), else 0
	int n_partials;  // accumulator copies = worker warps
	int n_ctas;
	int num_sms;
	int nz;
	int n_side;
	int lut_hi0, lut_shift, lut_n;
	int max_tasks;
	LutEntry lut[LUT_SIZE];
};

struct TiledArgs {
	DevParams P;
	const Cand *cand;
	const int32_t *cand_jk;
	const int64_t *cell_start;
	const CellInfo *cinfo;
	const ColInfo *colinfo;  // (r, mu_r) kernel only
	const int32_t *colreg;   // (r, mu_r) kernel only: [column][region] the one label of the region's candidates, -1 several, -2 none
	const double *slab_lo, *slab_hi;
	const Prim *prim;
	const int32_t *task_col;
	const int64_t *task_first;
	const int32_t *task_n;
	const int32_t *task_slab;  // [2 * task]: first and one-past-last slab of the task
	const unsigned long long *task_cost, *task_cum;
	const int32_t *n_tasks;
	const LutEntry *lut;
	int lut_hi0, lut_shift, lut_n;
	Accum A;
	int nz, n_side, n_workers, shard_index, shard_count, max_tasks;
	int w_r, ratio, hsplit, n_lr;  // (r, mu_r) kernel (ratio, n_lr: also the row-streaming (r_p, Pi) kernel)
	const double *vlo, *vhi;       // row-streaming (r_p, Pi) kernel: envelopes of the v coordinates per column index cv
	int ch_sym;                    // symmetric (r, mu_r) variant: candidates per staged chunk
	int *flags;
};

// defined in mia_tiled_rmu.cuh
inline bool rmu_supported(const mia_params *p, int &w_r);
inline size_t tiled_rmu_smem_bytes(bool unit_w, bool sig);
inline int launch_rmu(const TiledArgs &a, bool unit_w, bool los2, bool sig, int n_ctas, size_t smem, cudaStream_t st);
inline int launch_rmu_sym(const TiledArgs &a, bool unit_w, bool los2, int n_ctas, size_t smem, cudaStream_t st);
inline size_t tiled_rmu_sym_smem_bytes(bool unit_w, int ns, int ch);
inline int rmu_sym_chunk(bool unit_w, int ns);
inline bool plan_rmu_grid(const mia_params *p, int n_side, TiledConfig &cfg, int &nc, int &nz, int &k);
// defined in mia_tiled_rppi2.cuh
struct TiledWorkspace;
inline bool plan_rppi2_grid(const mia_params *p, int n_side, TiledConfig &cfg, int &nc, int nz, int &k);
inline size_t tiled_rppi2_smem_bytes(bool unit_w, bool sig);
inline int rppi2_prepare(const TiledConfig &cfg, const GridDims &g, const TiledWorkspace &w, cudaStream_t st);
inline int rppi2_fill_tasks(const TiledArgs &a, const int64_t *prim_cell_start, const int64_t *cell_start, const int32_t *task_off,
							int ncol_s, int nzs, int k, int split, int sym, int32_t *task_col, int64_t *task_first, int32_t *task_n,
							int32_t *task_slab, unsigned long long *task_cost, int32_t *n_tasks, cudaStream_t st);
inline int launch_rppi2(const TiledArgs &a, bool unit_w, bool sig, int n_ctas, size_t smem, cudaStream_t st);
// defined in mia_tiled_rppi2s.cuh
inline bool rppi2s_supported(int ncu, int k, int ratio);
inline size_t tiled_rppi2s_smem_bytes(bool unit_w);
inline int launch_rppi2s(const TiledArgs &a, bool unit_w, int n_ctas, size_t smem, cudaStream_t st);
inline int rmu_fill_tasks(const TiledArgs &a, const int64_t *prim_cell_start, const int64_t *cell_start, const int32_t *task_off,
						  int ncol_s, int nzs, int k, int split, int sym, int32_t *task_col, int64_t *task_first, int32_t *task_n,
						  int32_t *task_slab, unsigned long long *task_cost, int32_t *n_tasks, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------------------
// planning (host)
// ------------------------------------------------------------------------------------------------------------------
inline int env_int(const char *name, int dflt) {
	const char *v = getenv(name);
	return (v && *v) ? atoi(v) : dflt;
}

inline int hi_word(double x) {
	long long b;
	memcpy(&b, &x, 8);
	return (int)(b >> 32);
}

inline double from_hi_lo(int hi, unsigned lo) {
	long long b = ((long long)hi << 32) | lo;
	double x;
	memcpy(&x, &b, 8);
	return x;
}

// Look-up table over the high word of s = r_p^2: entry e covers hi in [hi0 + e*2^shift, hi0 + (e+1)*2^shift).
inline bool build_lut(const mia_params *p, TiledConfig &cfg) {
	const double *thr = p->r2_thr_host;
	const int n_r = p->n_r;
	if (!(thr[0] > 0.0) || !(thr[n_r] > thr[0]) || !std::isfinite(thr[n_r])) return false;
	const int hi0 = hi_word(thr[0]), hi1 = hi_word(thr[n_r]);
	int shift = 0;
	while ((((long long)hi1 - hi0) >> shift) + 1 > LUT_SIZE) shift++;
	const int n = (int)((((long long)hi1 - hi0) >> shift) + 1);
	for (int e = 0; e < n; e++) {
		const double lowest = from_hi_lo(hi0 + (e << shift), 0u);
		const double next_lowest = from_hi_lo(hi0 + ((e + 1) << shift), 0u);
		int base = 0, inside = 0;
		double t_in = INFINITY;
		for (int b = 1; b < n_r; b++) {
			if (thr[b] <= lowest) base++;
			else if (thr[b] < next_lowest) {
				inside++;
				t_in = thr[b];
			}
		}
		if (inside > 1) return false;  // bins finer than the table: leave it to the general kernel
		cfg.lut[e].thr = t_in;
		cfg.lut[e].base = base;
		cfg.lut[e].pad = 0;
	}
	for (int e = n; e < LUT_SIZE; e++) cfg.lut[e] = LutEntry{INFINITY, n_r - 1, 0};
	cfg.lut_hi0 = hi0;
	cfg.lut_shift = shift;
	cfg.lut_n = n;
	return true;
}

inline size_t tiled_smem_bytes(bool unit_w) {
	const size_t fixed = sizeof(Cand) * TW * STAGES * CH + sizeof(int) * TW * MAX_NEIGH + 256 + 768;
	const size_t per_slot = (size_t)TP * (8 + 8 + 4 + (unit_w ? 0 : 8));
	return fixed + per_slot * NSLOT;
}

inline bool plan_tiled(const mia_params *p, int64_t nD, int64_t nS, GridDims &g, int &ku, int &kv, int &kl,
					   TiledConfig &cfg) {
	const double L = p->boxsize, reach = p->r_search * (1.0 + 1e-6);
	int n_side = 1;
	while ((n_side + 1) * (n_side + 1) * (n_side + 1) <= (p->num_jk > 0 ? p->num_jk : 1)) n_side++;
	cfg.n_side = n_side;
	cfg.geom = p->geometry;
	cfg.w_r = 0;
	cfg.ratio = 1;
	cfg.hsplit = 1;
	cfg.n_lr = 1;
	cfg.v2 = 0;
	cfg.sym_ok = 0;
	cfg.sym = 0;
	cfg.sig = p->variance ? 1 : 0;
	// columns: about a quarter of the search radius wide
	int nc = (int)floor(L / (reach / 4.0));
	if (nc > 2048) nc = 2048;
	if (nc < 1) nc = 1;
	int nz;
	if (p->geometry == MIA_GEOM_RPPI) {
		// slabs thinner than the narrowest Pi bin
		double dmin = INFINITY;
		for (int b = 0; b < p->n_2; b++) {
			const double lo = p->thr2_host[b], hi = p->thr2_host[b + 1];
			if (!std::isfinite(lo) || !std::isfinite(hi) || !(hi > lo)) return false;
			dmin = fmin(dmin, hi - lo);
		}
		const double nz_d = floor(L / (dmin * (1.0 - 1e-9))) + 1.0;
		if (!(nz_d >= 1.0) || nz_d > 512.0) return false;
		nz = (int)nz_d;
		if (!build_lut(p, cfg)) return false;
		// align columns and slabs with the jackknife sub-boxes: a cell then carries one label (no label runs to scan, fewer
		// and larger chunks, fewer flushes)
		{
			const char *ev = getenv("MIA_RPPI_ALIGN");
			const int mode = ev ? atoi(ev) : MIA_RPPI_ALIGN;
			if (mode > 0 && n_side > 1 && nc >= 4 * n_side) {
				nc = (mode == 1) ? nc / n_side * n_side : (nc + n_side - 1) / n_side * n_side;
				const int nz2 = (nz + n_side - 1) / n_side * n_side;
				if (nz2 <= 512) nz = nz2;
			}
		}
	} else {
		// (r, mu_r): 3-D search with cubic cells aligned with the jackknife sub-boxes (a cell then carries one label)
		if (!rmu_supported(p, cfg.w_r)) return false;
	}
	int k;
	if (p->geometry == MIA_GEOM_RMU) {
		if (!plan_rmu_grid(p, n_side, cfg, nc, nz, k)) return false;
	} else {
		// Row-streaming kernel (mia_tiled_rppi2.cuh) or cell-by-cell kernel?  Measured (profiles/r01_tuning.md): rows win when
		// a (row, slab) range holds many candidates -- dense catalogues, slabs not too thin; 1 = decide by that estimate,
		// 0 = never, 2 = always.
		const char *ev = getenv("MIA_RPPI_V2");
		int v2 = ev ? atoi(ev) : MIA_RPPI_V2;
		int div2 = 0;
		bool thin_auto = false;
		if (cfg.sig && v2 == 0) v2 = 1;
		const bool must_rows = cfg.sig;  // only the row-streaming kernel has the variance variant
		if (v2 == 1) {
			const double rho = (double)nD / (L * L * L);
			auto piece = [&](int d) { return rho * (reach / d) * (1.6 * reach) * (L / nz); };  // candidates per streamed range
			// thin slabs (more than 16): the ordered rows kernel loses to the cell-by-cell kernel (196 vs 171 ms, cfg2 with 8 x 20
			// bins), the SYMMETRIC rows kernel wins (153 ms) -- so rows are planned when the two samples have the same size (an
			// auto-correlation, as far as the plan can know: the workspace layout must not depend on pointer identity) and the
			// grid admits the half-space rule (checked below)
			thin_auto = nz > 16 && nz <= 32 && nD == nS && env_int("MIA_SYM", 1) != 0 && p->kernel != MIA_KERNEL_TILED_ORDERED;
			// cells per r_max: the finest of 26 / 22 / 18 / 14 / 10 that still leaves ~110 candidates per streamed range (measured,
			// profiles/r02_tuning.md: cfg2 95.9 / 93.2 / 94.9 ms at 10 / 14 / 18; cfg4 3650 / 3419 / 3323 / 3247 / 3230 / 3250 ms at
			// 10 / 14 / 18 / 22 / 26 / 30): finer cells cull better and fill the 32 lanes of a round with rows
			if (nz > 16 && !thin_auto) v2 = 0;
			else if (thin_auto) {
				if (piece(10) >= 40.0) div2 = 10;
				else if (piece(6) >= 15.0) div2 = 6;
				else v2 = 0;
			} else {
				for (int d : {26, 22, 18, 14}) {
					// (the ordered kernel streams 2k + 2 rows: beyond 14 they no longer fit one round of 32 lanes; it measured
					// 114.4 / 113.4 / 118.5 ms at 10 / 14 / 18 on cfg2, so different sample sizes -- surely a cross-correlation -- stop at 14)
					if (d > 14 && nD != nS) continue;
					if (piece(d) >= 110.0) {
						div2 = d;
						break;
					}
				}
				if (!div2) {
					if (piece(10) >= 100.0) div2 = 10;
					else if (piece(6) >= 15.0) div2 = 6;
					else v2 = 0;
				}
			}
			if (v2 == 0 && must_rows) {
				v2 = 2;
				div2 = 6;
			}
		}
		if (v2) {  // row-streaming kernel: finer columns, coarser shape columns
			const int nc_cells = nc;
			cfg.w_r = div2;  // (r_p, Pi): carries the chosen cells per r_max to plan_rppi2_grid (0 = default)
			if (!plan_rppi2_grid(p, n_side, cfg, nc, nz, k)) return false;
			cfg.sym_ok = rppi2s_supported(nc, k, cfg.ratio) ? 1 : 0;
			if (thin_auto && !cfg.sym_ok && !must_rows) {  // no symmetric kernel after all: back to the cell-by-cell plan
				v2 = 0;
				cfg.v2 = 0;
				cfg.ratio = 1;
				cfg.n_lr = 1;
				cfg.w_r = 0;
				nc = nc_cells;
			}
		}
		if (!v2) {
			const double cs = L / nc;
			k = (int)ceil(reach / cs);
			if (k < 1) k = 1;
			const bool all_mode = (2 * k + 1 >= nc);
			if (all_mode ? ((long long)nc * nc > MAX_NEIGH) : ((2 * k + 1) * (2 * k + 1) > MAX_NEIGH)) return false;
		}
	}
	g.ncu = g.ncv = nc;
	g.ncl = nz;
	g.inv_cu = g.inv_cv = nc / L;
	g.inv_cl = nz / L;
	ku = kv = k;
	kl = nz;  // all slabs
	cfg.nz = nz;
	int dev = 0, sms = 148;
	if (cudaGetDevice(&dev) == cudaSuccess) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
	} else {
		cudaGetLastError();
	}
	cfg.num_sms = sms;
	cfg.n_ctas = sms * SLOTS_PER_SM;
	// worker slots (= accumulator copies) per SM: more slots = a shorter tail behind the last slot, more copies to zero and
	// reduce.  The cell-by-cell kernel assigns slots statically to its launched warps; the others hand them out dynamically.
	// Measured (profiles/r02_tuning.md, cfg2): 1x / 2x / 4x slots = 110.6 / 109.2 / 103.3 ms; small catalogues keep 1x (zeroing and
	// reducing 2 GB of copies costs 0.5 ms).
	{
		const int auto_mult = nS >= 500000 ? 4 : (nS >= 200000 ? 2 : 1);
		const int mult = env_int("MIA_SLOT_MULT", MIA_SLOT_MULT > 0 ? MIA_SLOT_MULT : auto_mult);
		cfg.n_partials = cfg.n_ctas * TW * ((cfg.v2 || p->geometry == MIA_GEOM_RMU) ? (mult < 1 ? 1 : mult) : 1);
	}
	{
		// one accumulator copy per worker slot: [2 * regions][bins] x 32 bytes.  Cap the total (default 4 GiB, MIA_ACC_CAP_MB)
		// by using fewer slots -- many jackknife regions x many bins would otherwise ask for tens of GB; slots are handed out
		// dynamically, so fewer slots only means fewer warps busy at the very end
		const size_t per_copy = (size_t)2 * (p->num_jk > 0 ? p->num_jk : 1) * p->n_r * p->n_2 * 32 + (size_t)p->n_r * p->n_2 * 8;
		const char *cap_env = getenv("MIA_ACC_CAP_MB");
		const size_t cap = (size_t)(cap_env && *cap_env ? atoll(cap_env) : 4096) << 20;
		size_t fit = cap / (per_copy ? per_copy : 1);
		if (fit < 64) fit = 64;
		if ((size_t)cfg.n_partials > fit) cfg.n_partials = (int)fit;
	}
	cfg.max_tasks = (int)((nS / (32 / cfg.hsplit) + (int64_t)nc * nc + 1) * MAX_SPLIT);
	return true;
}

struct TiledWorkspace {
	CellInfo *cinfo;
	ColInfo *colinfo;
	int32_t *colreg;
	double *vlo, *vhi;
	double *slab_lo, *slab_hi;
	int32_t *col_chunks, *task_off, *task_col, *task_n, *task_slab, *n_tasks;
	int64_t *task_first;
	unsigned long long *task_cost, *task_cum;
	LutEntry *lut;
	void *cub_tmp;
	size_t cub_bytes;
	size_t total;
};

inline size_t tiled_cub_bytes(int64_t ncol, int max_tasks) {
	size_t a = 0, b = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr, (int)(ncol + 1));
	cub::DeviceScan::InclusiveSum(nullptr, b, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
								  max_tasks);
	return a > b ? a : b;
}

inline TiledWorkspace carve_tiled(const TiledConfig &cfg, const GridDims &g, void *base) {
	TiledWorkspace w;
	size_t o = 0;
	unsigned char *b = (unsigned char *)base;
	auto take = [&](size_t bytes) {
		size_t at = o;
		o = (o + bytes + 255) / 256 * 256;
		return b ? (void *)(b + at) : (void *)nullptr;
	};
	const int64_t ncell = g.ncell(), ncol = (int64_t)g.ncu * g.ncv;
	w.cinfo = (CellInfo *)take(sizeof(CellInfo) * ncell);
	const int64_t n_info = cfg.geom == MIA_GEOM_RMU ? ncol : (cfg.v2 ? (int64_t)g.ncu * cfg.nz : 0);  // columns / (u row, slab)
	w.colinfo = (ColInfo *)take(sizeof(ColInfo) * n_info);
	w.colreg = (int32_t *)take(sizeof(int32_t) * n_info * cfg.n_lr);
	w.vlo = (double *)take(cfg.v2 ? sizeof(double) * g.ncv : 0);
	w.vhi = (double *)take(cfg.v2 ? sizeof(double) * g.ncv : 0);
	w.slab_lo = (double *)take(sizeof(double) * cfg.nz);
	w.slab_hi = (double *)take(sizeof(double) * cfg.nz);
	w.col_chunks = (int32_t *)take(sizeof(int32_t) * (ncol + 1));
	w.task_off = (int32_t *)take(sizeof(int32_t) * (ncol + 1));
	w.task_col = (int32_t *)take(sizeof(int32_t) * cfg.max_tasks);
	w.task_n = (int32_t *)take(sizeof(int32_t) * cfg.max_tasks);
	w.task_slab = (int32_t *)take(sizeof(int32_t) * 2 * cfg.max_tasks);
	w.n_tasks = (int32_t *)take(sizeof(int32_t) * 4);
	w.task_first = (int64_t *)take(sizeof(int64_t) * cfg.max_tasks);
	w.task_cost = (unsigned long long *)take(sizeof(unsigned long long) * cfg.max_tasks);
	w.task_cum = (unsigned long long *)take(sizeof(unsigned long long) * cfg.max_tasks);
	w.lut = (LutEntry *)take(sizeof(LutEntry) * LUT_SIZE);
	w.cub_bytes = tiled_cub_bytes(ncol, cfg.max_tasks);
	w.cub_tmp = take(w.cub_bytes);
	w.total = o;
	return w;
}

inline size_t tiled_workspace_bytes(const TiledConfig &cfg, const GridDims &g, int64_t, int64_t) {
	return carve_tiled(cfg, g, nullptr).total;
}

// ------------------------------------------------------------------------------------------------------------------
// preparation kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_cell_info(const unsigned char *__restrict__ cand, int stride, const int32_t *__restrict__ cand_jk,
							const int64_t *__restrict__ cell_start, int64_t ncell, int nz, CellInfo *__restrict__ info,
							unsigned long long *__restrict__ slab_lo, unsigned long long *__restrict__ slab_hi, int order = 0,
							int ncv = 1) {
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c >= ncell) return;
	const int64_t j0 = cell_start[c], j1 = cell_start[c + 1];
	CellInfo ci;
	ci.umin = ci.vmin = INFINITY;
	ci.umax = ci.vmax = -INFINITY;
	ci.label = -1;
	ci.nlab = 0;
	if (j1 > j0) {
		double lmin = INFINITY, lmax = -INFINITY;
		int prev = -1;
		for (int64_t j = j0; j < j1; j++) {
			// u, v, l lead every candidate record (Cand, CandSU, CandSW)
			const double2 uv = *reinterpret_cast<const double2 *>(cand + (size_t)j * stride);
			struct { double u, v, l; } q = {uv.x, uv.y, *reinterpret_cast<const double *>(cand + (size_t)j * stride + 16)};
			ci.umin = fmin(ci.umin, q.u);
			ci.umax = fmax(ci.umax, q.u);
			ci.vmin = fmin(ci.vmin, q.v);
			ci.vmax = fmax(ci.vmax, q.v);
			lmin = fmin(lmin, q.l);
			lmax = fmax(lmax, q.l);
			const int lab = cand_jk[j];
			if (j == j0) ci.label = lab;
			if (lab != prev) ci.nlab++;
			prev = lab;
		}
		const int s = order ? (int)((c / ncv) % nz) : (int)(c % nz);
		// l >= 0, so the bit pattern of (l + 0.0) is monotone in l
		atomicMin(&slab_lo[s], (unsigned long long)__double_as_longlong(lmin + 0.0));
		atomicMax(&slab_hi[s], (unsigned long long)__double_as_longlong(lmax + 0.0));
	}
	info[c] = ci;
}

// (r, mu_r) kernel: per column, the bounding box in the projected axes and, per line-of-sight region, the label its
// candidates share (-1: several labels, -2: no candidates).  Thread 0 also turns the per-slab coordinate bounds into
// envelopes (slab_lo[s] = smallest coordinate in slabs >= s, slab_hi[s] = largest in slabs <= s), which are valid
// bounds for any run of slabs, empty ones included.
__global__ void k_col_info(const CellInfo *__restrict__ cinfo, int64_t ncol, int nz, int n_lr, ColInfo *__restrict__ out,
						   int32_t *__restrict__ colreg, double *__restrict__ slab_lo, double *__restrict__ slab_hi) {
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c == 0) {
		double cur = INFINITY;
		for (int s = nz - 1; s >= 0; s--) {
			const double v = slab_lo[s];
			if (v == v) cur = fmin(cur, v);  // empty slabs still hold the NaN fill pattern
			slab_lo[s] = cur;
		}
		cur = -INFINITY;
		for (int s = 0; s < nz; s++) {
			const double lo = slab_lo[s], v = slab_hi[s];
			(void)lo;
			cur = fmax(cur, v);  // empty slabs hold 0.0 <= every coordinate
			slab_hi[s] = cur;
		}
	}
	if (c >= ncol) return;
	ColInfo o;
	o.umin = o.vmin = INFINITY;
	o.umax = o.vmax = -INFINITY;
	const int per = nz / n_lr;
	for (int r = 0; r < n_lr; r++) {
		int lab = -2;
		const int s1 = (r == n_lr - 1) ? nz : (r + 1) * per;
		for (int s = r * per; s < s1; s++) {
			const CellInfo ci = cinfo[c * nz + s];
			if (ci.nlab == 0) continue;
			o.umin = fmin(o.umin, ci.umin);
			o.umax = fmax(o.umax, ci.umax);
			o.vmin = fmin(o.vmin, ci.vmin);
			o.vmax = fmax(o.vmax, ci.vmax);
			if (ci.nlab > 1) lab = -1;
			else if (lab == -2) lab = ci.label;
			else if (lab != ci.label) lab = -1;
		}
		colreg[c * n_lr + r] = lab;
	}
	out[c] = o;
}

__global__ void k_col_chunks(const int64_t *__restrict__ prim_cell_start, int64_t ncol, int nzs, int split,
							 int32_t *__restrict__ col_chunks, int spt = 32) {
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c > ncol) return;
	if (c == ncol) {
		col_chunks[c] = 0;
		return;
	}
	const int64_t n = prim_cell_start[(c + 1) * nzs] - prim_cell_start[c * nzs];  // nzs = cells of the shape sort per column
	col_chunks[c] = (int32_t)((n + spt - 1) / spt) * split;
}

__device__ __forceinline__ bool neighbour_offset_ok(int ou, int ov, double cs, double reach) {
	const double mu = (abs(ou) > 1) ? (double)(abs(ou) - 1) : 0.0, mv = (abs(ov) > 1) ? (double)(abs(ov) - 1) : 0.0;
	return (mu * mu + mv * mv) * cs * cs * (1.0 - 1e-6) < reach * reach;
}

// One warp task = up to 32 consecutive shape galaxies of one column; cost = shapes x candidates in reach.
__global__ void k_fill_tasks(const int64_t *__restrict__ prim_cell_start, const int64_t *__restrict__ cell_start,
							 const int32_t *__restrict__ task_off, int ncu, int ncv, int nz, int nzs, int split, int k,
							 int periodic, double cs, double reach, int32_t *__restrict__ task_col,
							 int64_t *__restrict__ task_first, int32_t *__restrict__ task_n, int32_t *__restrict__ task_slab,
							 unsigned long long *__restrict__ task_cost, int32_t *__restrict__ n_tasks) {
	const int64_t ncol = (int64_t)ncu * ncv;
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c == 0) n_tasks[0] = task_off[ncol];
	if (c >= ncol) return;
	const int64_t p0 = prim_cell_start[c * nzs], p1 = prim_cell_start[(c + 1) * nzs];
	if (p1 <= p0) return;
	const int cu = (int)(c / ncv), cv = (int)(c % ncv);
	const bool all_u = 2 * k + 1 >= ncu, all_v = 2 * k + 1 >= ncv;
	unsigned long long W = 0;
	for (int iu = all_u ? 0 : -k; iu <= (all_u ? ncu - 1 : k); iu++) {
		int nu = all_u ? iu : cu + iu;
		if (nu < 0 || nu >= ncu) {
			if (!periodic) continue;
			nu = (nu + ncu) % ncu;
		}
		for (int iv = all_v ? 0 : -k; iv <= (all_v ? ncv - 1 : k); iv++) {
			int nv = all_v ? iv : cv + iv;
			if (nv < 0 || nv >= ncv) {
				if (!periodic) continue;
				nv = (nv + ncv) % ncv;
			}
			if (!all_u && !all_v && !neighbour_offset_ok(iu, iv, cs, reach)) continue;
			const int64_t nc = (int64_t)nu * ncv + nv;
			W += (unsigned long long)(cell_start[(nc + 1) * nz] - cell_start[nc * nz]);
		}
	}
	int t = task_off[c];
	for (int64_t p = p0; p < p1; p += 32) {
		const int n = (int)((p1 - p < 32) ? (p1 - p) : 32);
		for (int part = 0; part < split; part++, t++) {  // the same 32 shapes against consecutive ranges of slabs
			task_col[t] = (int32_t)c;
			task_first[t] = p;
			task_n[t] = n;
			task_slab[2 * t] = (int)((long long)nz * part / split);
			task_slab[2 * t + 1] = (int)((long long)nz * (part + 1) / split);
			task_cost[t] = (unsigned long long)n * W / (unsigned long long)split + 1ull;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// device helpers: mbarrier + 1-D bulk (TMA) copy
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"MIA_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra MIA_DONE;\n"
		"bra MIA_WAIT;\n"
		"MIA_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}
// global -> shared bulk copy (UBLKCP), completion signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					 smem_u32(dst)),
				 "l"(src), "r"(bytes), "r"(smem_u32(bar))
				 : "memory");
}

__device__ __forceinline__ double fast_rcp(double x) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double e = fma(-x, y, 1.0);
	y = fma(y, e, y);
	e = fma(-x, y, 1.0);
	return fma(y, e, y);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
	return x;
}

// Worker slots are handed out dynamically: a warp takes the next unprocessed slot until none is left.  A slot's tasks (fixed by
// the prefix sum of estimated work) are accumulated in their fixed order into the SLOT's accumulator copy, so the result does
// not depend on which warp ran which slot; what changes is that a warp that finishes early picks up more work (no tail behind
// the slowest warp of a CTA, no dependence on how many CTAs are resident).
__device__ __forceinline__ int next_slot(int *counter) {
	int s = 0;
	if ((threadIdx.x & 31) == 0) s = atomicAdd(counter, 1);
	return __shfl_sync(0xffffffffu, s, 0);
}

// extended bin of x on the second axis: -1 below range, n_2 at/above the upper range edge
__device__ __forceinline__ int ebin2(double x, const double *thr2, int n_2) {
	int c = 0;
	for (int b = 0; b <= n_2; b++) c += (x >= thr2[b]) ? 1 : 0;
	return c - 1;
}

// the few scalars z_window needs (passed by value: a noinline function taking the kernel's parameter block by reference
// would force a 1 KB copy of it into every thread's local memory)
struct ZParams {
	const double *thr2;  // shared-memory copy of the second-axis thresholds
	double L, halfL;
	int n_2, periodic;
};

struct ZWindow {
	double t_split, t_lo, t_hi;
	double shift;  // 0, -L or +L: the periodic image every pair of this lane with this slab takes
	// slab straddling +-L/2 for this lane: a pair takes the image `wadd` iff (d > wthr) != wflip, else `shift` (= 0).
	// "d < -L/2" is written as !(d > nextbelow(-L/2)), so one comparison serves both directions.
	double wthr, wadd;
	bool wflip;
	int b0, b1;
	bool gen, dead, err;
};

// Which (at most two) Pi bins can a pair fall in whose raw line-of-sight separation (before the periodic shift) lies in
// [lo, hi]?  (hi - lo is smaller than the narrowest Pi bin.)
__device__ __noinline__ ZWindow z_window_sep(double lo, double hi, const ZParams P) {
	ZWindow w;
	w.t_split = INFINITY;
	w.t_lo = -INFINITY;
	w.t_hi = INFINITY;
	w.shift = 0.0;
	w.wthr = INFINITY;
	w.wadd = 0.0;
	w.wflip = false;
	w.b0 = w.b1 = -1;
	w.gen = false;
	w.dead = false;
	w.err = false;
	const int n2 = P.n_2;
	bool straddle = false;
	if (P.periodic && !(lo >= -P.halfL && hi <= P.halfL)) {
		if (lo > P.halfL) {  // every pair wraps down: sep -= L (measure_w_box_jk.py:403)
			lo = __dsub_rn(lo, P.L);
			hi = __dsub_rn(hi, P.L);
			w.shift = -P.L;
		} else if (hi < -P.halfL) {  // every pair wraps up (:404)
			lo = __dadd_rn(lo, P.L);
			hi = __dadd_rn(hi, P.L);
			w.shift = P.L;
		} else {
			straddle = true;
			w.gen = true;
		}
	}
	if (!straddle) {
		const int ea = ebin2(lo, P.thr2, P.n_2), eb = ebin2(hi, P.thr2, P.n_2);
		if (eb - ea > 1) w.err = true;
		if (ea >= 0 && ea < n2) w.b0 = ea;
		if (eb != ea) {
			if (eb >= 0 && eb < n2) w.b1 = eb;
			w.t_split = P.thr2[eb <= n2 ? (eb < 0 ? 0 : eb) : n2];
		}
		if (ea < 0) {
			w.t_lo = P.thr2[0];
			w.gen = true;
		}
		if (eb >= n2) {
			w.t_hi = P.thr2[n2];
			w.gen = true;
		}
	} else {
		// part A: values close to +L/2, part B: values close to -L/2 (one of them wrapped)
		double a_lo, b_hi;
		if (hi > P.halfL) {  // pairs with d > L/2 wrap down
			a_lo = lo;
			b_hi = __dsub_rn(hi, P.L);
			w.wthr = P.halfL;
			w.wadd = -P.L;
			w.wflip = false;
		} else {  // pairs with d < -L/2 wrap up
			a_lo = __dadd_rn(lo, P.L);
			b_hi = hi;
			w.wthr = __longlong_as_double(__double_as_longlong(-P.halfL) + 1);  // next double below -L/2
			w.wadd = P.L;
			w.wflip = true;
		}
		const int ea_a = ebin2(a_lo, P.thr2, P.n_2), eb_a = ebin2(P.halfL, P.thr2, P.n_2);
		const int ea_b = ebin2(-P.halfL, P.thr2, P.n_2), eb_b = ebin2(b_hi, P.thr2, P.n_2);
		int na = 0, nbb = 0, va = -1, vb = -1;
		for (int e = ea_a; e <= eb_a; e++)
			if (e >= 0 && e < n2) {
				na++;
				va = e;
			}
		for (int e = ea_b; e <= eb_b; e++)
			if (e >= 0 && e < n2) {
				nbb++;
				vb = e;
			}
		if (na > 1 || nbb > 1) w.err = true;
		w.b0 = vb;
		w.b1 = va;
		w.t_split = 0.0;
		w.t_lo = P.thr2[0];
		w.t_hi = P.thr2[n2];
		if (w.b0 == w.b1) {
			w.b1 = -1;
			w.t_split = INFINITY;
		}
	}
	w.dead = (w.b0 < 0 && w.b1 < 0);
	return w;
}

// The same for one shape galaxy at line-of-sight coordinate pl against the candidates of a slab [zlo, zhi].
__device__ __forceinline__ ZWindow z_window(double pl, double zlo, double zhi, const ZParams P) {
	return z_window_sep(__dsub_rn(pl, zhi), __dsub_rn(pl, zlo), P);  // fl(s - c) is monotone in c
}

// ------------------------------------------------------------------------------------------------------------------
// explicit shared-memory access (32-bit shared addresses; keeps nvcc from re-deriving the shared window per access)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds_v2(double &a, double &b, uint32_t addr) {
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_v2(uint32_t addr, double a, double b) {
	asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
	double a;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
	return a;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double a) {
	asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(uint32_t addr) {
	unsigned a;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr));
	return a;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, unsigned a) {
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void lds_lut(double &thr, int &base, uint32_t addr) {
	long long a, b;
	asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
	thr = __longlong_as_double(a);
	base = (int)b;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// predicated shared stores: the pair loop is branch-free so that nvcc can interleave the arithmetic of consecutive
// candidates (the FP64 pipe has ~8-cycle dependent-issue latency and only 3 warps per scheduler fit)
__device__ __forceinline__ void sts_v2_if(bool p, uint32_t addr, double a, double b) {
	asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(addr), "d"(a), "d"(b),
				 "r"((unsigned)p)
				 : "memory");
}
__device__ __forceinline__ void sts_f64_if(bool p, uint32_t addr, double a) {
	asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(addr), "d"(a), "r"((unsigned)p)
				 : "memory");
}
__device__ __forceinline__ void sts_u32_if(bool p, uint32_t addr, unsigned a) {
	asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u32 [%0], %1; }" ::"r"(addr), "r"(a), "r"((unsigned)p)
				 : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// the pair kernel
// ------------------------------------------------------------------------------------------------------------------
// Shared-memory addresses (bytes, shared window) of the thread-private accumulators:
//   a2 : [slot][thread] double2 {sum e+ , sum ex}     ac : [slot][thread] u32 pair count     aw : [slot][thread] sum w_D
//   av : [slot][thread] sum (w_D e+)^2 (SIG variants: the brute variants' `variance`, measure_w_box_jk.py:196)
struct PrivAcc {
	uint32_t a2, ac, aw, av;
};

// One accumulation window of r bins [ra, ra + W_R): its squared-separation limits and interior thresholds.
struct RWindow {
	double lo, hi;        // pairs with lo <= r_p^2 < hi belong to the window
	double thr[W_R - 1];  // interior thresholds (+inf when the window has fewer bins): r bin = ra + #{thr <= r_p^2}
	int ra;
};

// One staged chunk (n candidates at shared address cb) against this thread's shape galaxy.
//   XYS : the projected separations take the lane-constant periodic image (su, sv) (column pair across the box edge)
//   ZW  : the line-of-sight separation is wrapped per pair with the lane's one-sided rule (slab straddling +-L/2) and
//         range-checked against the Pi axis; otherwise only the lane-constant image shift is added
//   GEN : everything compared and wrapped per pair (tiny boxes, where a column pair can straddle +-L/2 as well)
// All of these are exactly the reference's `sep -= L` / `sep += L` (measure_w_box_jk.py:403-404): x + 0.0 == x.
// Returns true when some candidate has |cos| within 1e-11 of 1 for this lane: those pairs are NOT accumulated here
// but re-evaluated with the reference's exact operation sequence by slow_pairs() (its NaN rule, :416-417).
template <bool UNITW, bool XYS, bool ZW, bool GEN, bool SIG = false>
__device__ __forceinline__ bool pair_loop(uint32_t cb, int n, int periodic, double L, double halfL, double pu, double pv,
										  double pl, double a0, double a1, double su, double sv, const RWindow &rw,
										  double hi_lane, const ZWindow &zw, const PrivAcc &acc) {
	// periodic image of one separation, branch-free and exactly the reference's two conditional shifts (:403-404):
	// |d| > L/2  =>  d -= copysign(L, d)   (after the first shift the second condition can no longer hold)
	auto wrap = [&](double d) {
		const double sl = __hiloint2double(__double2hiint(L) | (__double2hiint(d) & 0x80000000), __double2loint(L));
		return (fabs(d) > halfL) ? __dsub_rn(d, sl) : d;  // sl = copysign(L, d)
	};
	bool lane_susp = false;
	// current candidate, the next one and the one after it (prefetch distance 2).  The prefetch may run up to two
	// candidates past the end of the chunk: that is still inside this CTA's shared memory and the values are never used.
	double cu, cv, cl, cw, mu_, mv_, ml_, mw_;
	lds_v2(cu, cv, cb);
	lds_v2(cl, cw, cb + 16);
	lds_v2(mu_, mv_, cb + (uint32_t)sizeof(Cand));
	lds_v2(ml_, mw_, cb + (uint32_t)sizeof(Cand) + 16);
	uint32_t na = cb + 2u * (uint32_t)sizeof(Cand);
#if MIA_SWP
	bool p_ok = false;  // the candidate whose accumulation is still pending
	uint32_t p_so = 0u;
	unsigned p_c0 = 0u;
	double p_s0 = 0.0, p_s1 = 0.0, p_gp = 0.0, p_gc = 0.0, p_sw = 0.0, p_cw = 0.0;
#endif
	MIA_UNROLL_PRAGMA(MIA_UNROLL)
	for (int j = 0; j < n; j++) {
		double nu, nv, nl, nw;
		lds_v2(nu, nv, na);
		lds_v2(nl, nw, na + 16);
		na += (uint32_t)sizeof(Cand);
		double du = __dsub_rn(pu, cu), dv = __dsub_rn(pv, cv), dz = __dsub_rn(pl, cl);  // shape minus position, :401
		if (GEN) {
			if (periodic) {
				du = wrap(du);
				dv = wrap(dv);
				dz = wrap(dz);
			}
		} else {
			if (XYS) {
				du = __dadd_rn(du, su);
				dv = __dadd_rn(dv, sv);
			}
			if (ZW) dz = __dadd_rn(dz, ((dz > zw.wthr) != zw.wflip) ? zw.wadd : zw.shift);
			else dz = __dadd_rn(dz, zw.shift);
		}
		const double r2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));  // :407 (before the sqrt)
		bool ok = (r2 >= rw.lo) && (r2 < hi_lane);
		if (ZW || GEN) ok = ok && (dz >= zw.t_lo) && (dz < zw.t_hi);
		int slot = (dz >= zw.t_split) ? 1 : 0;
#pragma unroll
		for (int k = 0; k < W_R - 1; k++) slot += (r2 >= rw.thr[k]) ? 2 : 0;
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
#if !MIA_SWP
		// private slots: loads first, the arithmetic below hides their latency
		double s0, s1, sw = 0.0, sq = 0.0;
		lds_v2(s0, s1, acc.a2 + so * 16u);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) sw = lds_f64(acc.aw + so * 8u);
		if (SIG) sq = lds_f64(acc.av + so * 8u);
#endif
		const double cr = fma(du, a0, __dmul_rn(dv, a1));   // r_p cos(phi)
		const double sr = fma(du, a1, -__dmul_rn(dv, a0));  // r_p sin(phi) (sign irrelevant)
		// 2 / r2: 20-bit hardware seed, one cubically convergent step y(1 + e + e^2) (relative error ~1e-17; a plain
		// Newton step leaves a one-signed 1e-12 bias that survives the cancellation in the S+D sums), then doubling by
		// an exponent increment
		double y;
		asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
		{
			const double e = fma(-r2, y, 1.0);
			y = fma(y, fma(e, e, e), y);
		}
		const double inv2 = __hiloint2double(__double2hiint(y) + 0x00100000, __double2loint(y));
		double gp = fma(cr * cr, inv2, -1.0);  // cos 2phi = 2 cos^2 - 1
		double gc = (cr * fabs(sr)) * inv2;    // sin 2phi = 2 cos phi |sin phi|   (phi in [0, pi])
		const bool susp = ok && (gp >= 1.0 - 1e-11);  // |cos| ~ 1: the reference's NaN rule may apply -> exact path
		lane_susp = lane_susp || susp;
		ok = ok && !susp;
#if MIA_SWP
		// software pipeline of the accumulation: retire the PREVIOUS candidate (its slot was loaded one iteration ago, so
		// the load latency is covered by a whole iteration of arithmetic), then load this candidate's slot.  Stores and
		// loads are volatile and stay in this order, so two consecutive candidates in the same slot are handled correctly.
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64_if(p_ok, acc.aw + p_so * 8u, p_sw + p_cw);
		}
		sts_v2_if(p_ok, acc.a2 + p_so * 16u, p_s0 + p_gp, p_s1 + p_gc);
		sts_u32_if(p_ok, acc.ac + p_so * 4u, p_c0 + 1u);
		lds_v2(p_s0, p_s1, acc.a2 + so * 16u);
		p_c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) p_sw = lds_f64(acc.aw + so * 8u);
		p_ok = ok;
		p_so = so;
		p_gp = gp;
		p_gc = gc;
		p_cw = cw;
#else
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64_if(ok, acc.aw + so * 8u, sw + cw);
		}
		if (SIG) sts_f64_if(ok, acc.av + so * 8u, fma(gp, gp, sq));
		sts_v2_if(ok, acc.a2 + so * 16u, s0 + gp, s1 + gc);
		sts_u32_if(ok, acc.ac + so * 4u, c0 + 1u);
#endif
		cu = mu_;
		cv = mv_;
		cl = ml_;
		cw = mw_;
		mu_ = nu;
		mv_ = nv;
		ml_ = nl;
		mw_ = nw;
	}
#if MIA_SWP
	if (!UNITW) sts_f64_if(p_ok, acc.aw + p_so * 8u, p_sw + p_cw);
	sts_v2_if(p_ok, acc.a2 + p_so * 16u, p_s0 + p_gp, p_s1 + p_gc);
	sts_u32_if(p_ok, acc.ac + p_so * 4u, p_c0 + 1u);
#endif
	return lane_susp;
}

// Rare path: rescan the chunk for the pairs pair_loop skipped (|cos| ~ 1), with the reference's exact operation
// sequence for the separation, cos and its NaN rule.
template <bool UNITW, bool SIG = false>
__device__ __noinline__ void slow_pairs(bool lane_susp, uint32_t cb, int n, int periodic, double L, double halfL, double pu,
										double pv, double pl, double a0, double a1, const RWindow rw, double w_hi, double t_lo,
										double t_hi, double t_split, PrivAcc acc, unsigned long long &nan_pairs) {
	const double w_lo = rw.lo;
	if (!lane_susp) return;
	auto sep = [&](double s_, double c_) {  // measure_w_box_jk.py:401-404
		double d = __dsub_rn(s_, c_);
		if (periodic) {
			if (d > halfL) d = __dsub_rn(d, L);
			if (d < -halfL) d = __dadd_rn(d, L);
		}
		return d;
	};
	for (int j = 0; j < n; j++) {
		double cu, cv, cl, cw;
		lds_v2(cu, cv, cb + (uint32_t)j * (uint32_t)sizeof(Cand));
		lds_v2(cl, cw, cb + (uint32_t)j * (uint32_t)sizeof(Cand) + 16);
		const double du = sep(pu, cu), dv = sep(pv, cv), dz = sep(pl, cl);
		const double r2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
		if (!((r2 >= w_lo) && (r2 < w_hi) && (dz >= t_lo) && (dz < t_hi))) continue;
		// the same (approximate) test pair_loop used to skip the pair
		const double cr = fma(du, a0, __dmul_rn(dv, a1));
		double y;
		asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
		{
			const double e = fma(-r2, y, 1.0);
			y = fma(y, fma(e, e, e), y);
		}
		const double inv2 = __hiloint2double(__double2hiint(y) + 0x00100000, __double2loint(y));
		if (!(fma(cr * cr, inv2, -1.0) >= 1.0 - 1e-11)) continue;
		const double rp = __dsqrt_rn(r2);
		const double c = __dadd_rn(__dmul_rn(__ddiv_rn(du, rp), a0), __dmul_rn(__ddiv_rn(dv, rp), a1));
		double gp = 0.0, gc = 0.0;
		if (fabs(c) <= 1.0) shape_projection(c, gp, gc);
		else nan_pairs++;
		int slot = (dz >= t_split) ? 1 : 0;
#pragma unroll
		for (int k = 0; k < W_R - 1; k++) slot += (r2 >= rw.thr[k]) ? 2 : 0;
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double s0, s1;
		lds_v2(s0, s1, acc.a2 + so * 16u);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64(acc.aw + so * 8u, lds_f64(acc.aw + so * 8u) + cw);
		}
		if (SIG) sts_f64(acc.av + so * 8u, fma(gp, gp, lds_f64(acc.av + so * 8u)));
		sts_v2(acc.a2 + so * 16u, s0 + gp, s1 + gc);
		sts_u32(acc.ac + so * 4u, c0 + 1u);
	}
}

// Flush: fixed-order warp reduction of the private slots into this warp's accumulator copy in HBM.
// Lanes are grouped by key = (jackknife label of the shape galaxy, its two Pi bins); groups are processed one after the
// other, every slot of a group is summed over the lanes with a butterfly, and lane s adds slot s to the A / B rows.
struct FlushCtx {
	unsigned long long *pcnt;
	double *pddw, *psp, *psc;
	double *pvar;  // [nb] of this warp's copy (SIG variants), else NULL
	int *flags;
	int n_2, nb, J, num_jk;
};

template <bool UNITW, bool SIG = false>
__device__ __noinline__ unsigned flush_slots(const FlushCtx &fc, PrivAcc acc, unsigned key, bool dead, double pe, double pw,
											 int ra, int rb, int jkD) {
	const int lane = threadIdx.x & 31;
	unsigned binned = 0;
	unsigned todo = __ballot_sync(0xffffffffu, !dead);
	while (todo) {
		const int leader = __ffs(todo) - 1;
		const unsigned k = __shfl_sync(0xffffffffu, key, leader);
		const unsigned grp = __ballot_sync(0xffffffffu, key == k) & todo;
		const bool in = (grp >> lane) & 1u;
		unsigned tot_cnt = 0;
		double tot_sp = 0.0, tot_sc = 0.0, tot_dw = 0.0, tot_sq = 0.0;
#pragma unroll 1
		for (int sl = 0; sl < NSLOT; sl++) {
			const uint32_t so = (uint32_t)sl * TP;
			const unsigned c = in ? lds_u32(acc.ac + so * 4u) : 0u;
			const unsigned csum = __reduce_add_sync(0xffffffffu, c);
			if (csum == 0u) continue;
			double v0 = 0.0, v1 = 0.0;
			if (in) lds_v2(v0, v1, acc.a2 + so * 16u);
			const double xs = warp_sum(v0 * pe);
			const double ys = warp_sum(v1 * pe);
			const double zs = UNITW ? (double)csum : warp_sum(in ? lds_f64(acc.aw + so * 8u) * pw : 0.0);
			const double qs = SIG ? warp_sum(in ? lds_f64(acc.av + so * 8u) * (pe * pe) : 0.0) : 0.0;
			if (lane == sl) {
				tot_cnt = csum;
				tot_sp = xs;
				tot_sc = ys;
				tot_dw = zs;
				tot_sq = qs;
			}
		}
		if (lane < NSLOT && tot_cnt) {
			const int kb0 = (int)((k >> 8) & 0xffu) - 1, kb1 = (int)(k & 0xffu) - 1, kjk = (int)(k >> 16);
			const int b2 = (lane & 1) ? kb1 : kb0;
			const int rbin = ra + (lane >> 1);
			if (b2 < 0 || rbin > rb) {
				atomicExch(&fc.flags[1], 1);
			} else {
				// rows A[label of the shape] and B[label of the chunk]; all loads are issued before the first store (one memory
				// latency instead of seven: the flush is latency-bound, ncu profiles/r02_*)
				const size_t bin = (size_t)rbin * fc.n_2 + b2;
				const size_t ia = (size_t)kjk * fc.nb + bin;
				const bool has_b = fc.num_jk > 0 && jkD != kjk;
				const size_t ib = has_b ? (size_t)(fc.J + jkD) * fc.nb + bin : ia;
				const unsigned long long c_a = fc.pcnt[ia], c_b = fc.pcnt[ib];
				const double d_a = fc.pddw[ia], p_a = fc.psp[ia], x_a = fc.psc[ia], d_b = fc.pddw[ib], p_b = fc.psp[ib];
				fc.pcnt[ia] = c_a + tot_cnt;
				fc.pddw[ia] = d_a + tot_dw;
				fc.psp[ia] = p_a + tot_sp;
				fc.psc[ia] = x_a + tot_sc;
				if (has_b) {
					fc.pcnt[ib] = c_b + tot_cnt;
					fc.pddw[ib] = d_b + tot_dw;
					fc.psp[ib] = p_b + tot_sp;
				}
				if (SIG) fc.pvar[bin] += tot_sq;
				binned += tot_cnt;
			}
		}
		todo &= ~grp;
	}
	__syncwarp();
#pragma unroll
	for (int sl = 0; sl < NSLOT; sl++) {
		sts_v2(acc.a2 + (uint32_t)sl * TP * 16u, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)sl * TP * 8u, 0.0);
		if (SIG) sts_f64(acc.av + (uint32_t)sl * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)sl * TP * 4u, 0u);
	}
	return binned;
}

// Neighbour columns of a task's column, ordered by jackknife (u, v) region so that candidate labels change rarely.
// Writes the list to nlist (per-warp shared scratch) and returns its length.
__device__ __noinline__ int build_neighbour_list(int *nlist, int col, int ncu, int ncv, int ku, int kv, int periodic,
												 int n_side, double cs, double reach) {
	const int lane = threadIdx.x & 31;
	const int cu0 = col / ncv, cv0 = col % ncv;
	const bool all_u = 2 * ku + 1 >= ncu, all_v = 2 * kv + 1 >= ncv;
	const int wu = all_u ? ncu : 2 * ku + 1, wv = all_v ? ncv : 2 * kv + 1;
	const int n_off = wu * wv, n_keys = n_side * n_side;
	int my_col[MAX_NEIGH / 32], my_key[MAX_NEIGH / 32];
#pragma unroll 1
	for (int r = 0; r < MAX_NEIGH / 32; r++) {
		const int o = r * 32 + lane;
		int c_ = -1, k_ = -1;
		if (o < n_off) {
			const int iu = o / wv, iv = o - iu * wv;
			int nu = all_u ? iu : cu0 + iu - ku, nv = all_v ? iv : cv0 + iv - kv;
			bool ok = true;
			if (nu < 0) {
				if (!periodic) ok = false;
				nu += ncu;
			} else if (nu >= ncu) {
				if (!periodic) ok = false;
				nu -= ncu;
			}
			if (nv < 0) {
				if (!periodic) ok = false;
				nv += ncv;
			} else if (nv >= ncv) {
				if (!periodic) ok = false;
				nv -= ncv;
			}
			if (ok && !all_u && !all_v && !neighbour_offset_ok(iu - ku, iv - kv, cs, reach)) ok = false;
			if (ok) {
				c_ = nu * ncv + nv;
				k_ = ((nu * n_side) / ncu) * n_side + (nv * n_side) / ncv;
			}
		}
		if (r == 0) { my_col[0] = c_; my_key[0] = k_; }
		if (r == 1) { my_col[1] = c_; my_key[1] = k_; }
		if (r == 2) { my_col[2] = c_; my_key[2] = k_; }
		if (r == 3) { my_col[3] = c_; my_key[3] = k_; }
	}
	__syncwarp();
	int nn = 0;
	for (int k = 0; k < n_keys; k++) {  // counting sort by region key, stable in offset order
#pragma unroll
		for (int r = 0; r < MAX_NEIGH / 32; r++) {
			const unsigned m = __ballot_sync(0xffffffffu, my_key[r] == k);
			if (my_key[r] == k) nlist[nn + __popc(m & ((1u << lane) - 1u))] = my_col[r];
			nn += __popc(m);
		}
	}
	__syncwarp();
	return nn;
}

struct Chunk {
	long long start;
	int n, label;
	int xy;  // 0: no periodic image in the projected axes, 1: lane-constant image shifts, 2: per-pair wrap needed
};

template <bool UNITW>
__global__ void __launch_bounds__(TP, MIA_MIN_CTAS) k_tiled_rppi(const TiledArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	const DevParams &P = a.P;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb = P.n_r * P.n_2;
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	const int periodic = P.periodic;
	const double L = P.L, halfL = P.halfL;

	// ---- shared memory carve-up ------------------------------------------------------------------------------------
	Cand *ring = reinterpret_cast<Cand *>(smem);  // [warp][stage][CH]: every warp runs its own double-buffered stream
	int *nlist_all = reinterpret_cast<int *>(smem + sizeof(Cand) * TW * STAGES * CH);
	uint64_t *full = reinterpret_cast<uint64_t *>(nlist_all + TW * MAX_NEIGH);  // [warp][stage]
	double *thr2_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(full) + 256);
	unsigned char *accbase = reinterpret_cast<unsigned char *>(thr2_s) + 768;
	const uint32_t acc_u32 = smem_u32(accbase);
	Cand *my_ring = ring + (size_t)warp * STAGES * CH;
	uint64_t *my_full = full + warp * STAGES;
	int *nlist = nlist_all + warp * MAX_NEIGH;
	const uint32_t my_ring_u32 = smem_u32(my_ring);
	PrivAcc acc;  // addresses of slot 0 of this thread
	acc.a2 = acc_u32 + (uint32_t)tid * 16u;
	acc.aw = acc_u32 + (uint32_t)NSLOT * TP * 16u + (uint32_t)tid * 8u;
	acc.ac = acc_u32 + (uint32_t)NSLOT * TP * (UNITW ? 16u : 24u) + (uint32_t)tid * 4u;
	acc.av = 0u;  // (the cell-by-cell kernel has no variance variant: such calls are planned onto the row-streaming kernel)

	if (tid == 0) {
		for (int s = 0; s < TW * STAGES; s++) mbar_init(&full[s], 1);
		mbar_fence_init();
		if (blockIdx.x == 0) a.A.stats[6] = (unsigned long long)a.n_tasks[0];
	}
	for (int e = tid; e <= P.n_2; e += blockDim.x) thr2_s[e] = P.thr2[e];
#pragma unroll
	for (int s = 0; s < NSLOT; s++) {
		sts_v2(acc.a2 + (uint32_t)s * TP * 16u, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)s * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)s * TP * 4u, 0u);
	}
	__syncthreads();  // the only CTA-wide synchronisation: from here on every warp works alone

	// ---- this warp's share of the tasks: worker slots of equal estimated work -------------------------------------------
	int task0 = 0, task1 = 0;
	{
		const int nt = a.n_tasks[0];
		if (nt > 0) {
			const double total2 = 2.0 * (double)a.task_cum[nt - 1];
			const int RG = a.shard_count * a.n_workers;
			const int mine = ((int)blockIdx.x * TW + warp) * a.shard_count + a.shard_index;  // slots interleaved across ranks
			auto slot_of = [&](int t) {
				const double mid2 = 2.0 * (double)a.task_cum[t] - (double)a.task_cost[t];
				int s = (int)(mid2 / total2 * (double)RG);
				return s < RG - 1 ? s : RG - 1;
			};
			auto lower = [&](int target) {  // first task with slot >= target
				int lo = 0, hi = nt;
				while (lo < hi) {
					const int mid = (lo + hi) >> 1;
					if (slot_of(mid) >= target) hi = mid;
					else lo = mid + 1;
				}
				return lo;
			};
			task0 = lower(mine);
			task1 = lower(mine + 1);
		}
	}

	// this warp's accumulator copy in HBM
	const size_t part = (size_t)(blockIdx.x * TW + warp) * (size_t)a.A.rows * nb;
	unsigned long long *pcnt = a.A.cnt + part;
	double *pddw = a.A.ddw + part, *psp = a.A.sp + part, *psc = a.A.sc + part;

	ZParams zp;
	zp.thr2 = thr2_s;
	zp.L = L;
	zp.halfL = halfL;
	zp.n_2 = P.n_2;
	zp.periodic = periodic;
	FlushCtx fc;
	fc.pcnt = pcnt;
	fc.pddw = pddw;
	fc.psp = psp;
	fc.psc = psc;
	fc.pvar = nullptr;
	fc.flags = a.flags;
	fc.n_2 = P.n_2;
	fc.nb = nb;
	fc.J = J;
	fc.num_jk = P.num_jk;
	uint32_t phase0 = 0u, phase1 = 0u;  // parity of the two stages of this warp's stream
	int st_issue = 0;                   // stage the next bulk copy goes to
	unsigned long long tested = 0, binned = 0, nan_pairs = 0;
	const double TN = P.r2_thr[P.n_r];
	const double cs = L / P.ncu, reach = sqrt(TN) * (1.0 + 1e-6);
	const int n_win = (P.n_r + W_R - 1) / W_R;

	for (int task = task0; task < task1; task++) {
		const int col = a.task_col[task];
		const int np = a.task_n[task];
		const bool active = lane < np;
		Prim p;
		if (active) {
			p = a.prim[a.task_first[task] + lane];
		} else {
			p.u = p.v = p.l = 0.0;
			p.w = 0.0;
			p.a0 = 1.0;
			p.a1 = 0.0;
			p.e = 0.0;
			p.jk = 0;
			p.orig = -1;
		}
		const double pe = p.w * p.e;

		// ---- neighbour columns of this task (ordered by jackknife (u, v) region) -----------------------------------
		const int nn = build_neighbour_list(nlist, col, P.ncu, P.ncv, P.ku, P.kv, periodic, a.n_side, cs, reach);

		const int slab0 = a.task_slab[2 * task], slab1 = a.task_slab[2 * task + 1];
		for (int s = slab0; s < slab1; s++) {
			const double zlo = a.slab_lo[s], zhi = a.slab_hi[s];
			if (!(zlo <= zhi)) continue;  // empty slab (uniform branch)

			ZWindow zw = z_window(p.l, zlo, zhi, zp);
			if (!active) zw.dead = true;
			if (active && zw.err) atomicExch(&a.flags[1], 1);
			if (!__any_sync(0xffffffffu, !zw.dead)) continue;
			const bool warp_zg = __any_sync(0xffffffffu, !zw.dead && zw.gen);
			const unsigned key = zw.dead ? 0xffffffffu
										 : (((unsigned)p.jk << 16) | ((unsigned)(zw.b0 + 1) << 8) | (unsigned)(zw.b1 + 1));

			for (int q = 0; q < n_win; q++) {
				// ---- accumulation window q: r bins [ra, rb] counted from the top ------------------------------------------
				const int rb = P.n_r - 1 - q * W_R, ra = (rb - W_R + 1 > 0) ? rb - W_R + 1 : 0;
				RWindow rw;
				rw.ra = ra;
				rw.lo = P.r2_thr[ra];
				rw.hi = P.r2_thr[rb + 1];
#pragma unroll
				for (int k = 0; k < W_R - 1; k++) rw.thr[k] = (ra + 1 + k <= rb) ? P.r2_thr[ra + 1 + k] : INFINITY;
				const double win_hi = rw.hi;
				const double hi_lane = zw.dead ? -1.0 : rw.hi;  // dead lanes never pass the range test

				auto flush = [&](int jkD) {
					binned += flush_slots<UNITW>(fc, acc, key, zw.dead, pe, p.w, ra, rb, jkD);
				};

				// ---- chunk generator over the neighbour cells of this slab (32 cell descriptors at a time, one per lane,
				// read through shuffles); a single call site keeps the kernel small enough for the instruction cache -------
				int g_base = 0;
				unsigned g_nonempty = 0u;
				long long d_start = 0;
				int d_n = 0, d_label = -1, d_nlab = 0;
				double d_umin = 0.0, d_umax = 0.0, d_vmin = 0.0, d_vmax = 0.0;
				long long g_pos = 0, g_run_end = 0, g_cell_end = 0;
				int g_label = -1, g_nlab = 1;
				int g_xy = 0;
				double g_su = 0.0, g_sv = 0.0;  // this lane's projected image shifts for the current cell
				auto next_chunk = [&](Chunk &c) -> bool {
					for (;;) {
						if (g_pos < g_run_end) {
							// split the rest of the run into equal chunks (66 -> 22 + 22 + 22, not 32 + 32 + 2)
							const int rest = (int)(g_run_end - g_pos), nch = (rest + CH - 1) / CH;
							c.start = g_pos;
							c.n = (rest + nch - 1) / nch;
							c.label = g_label;
							c.xy = g_xy;
							g_pos += c.n;
							return true;
						}
						if (g_pos < g_cell_end) {  // next label run of a cell cut by a jackknife face
							g_label = a.cand_jk[g_pos];
							long long qq = g_pos + 1;
							while (qq < g_cell_end && a.cand_jk[qq] == g_label) qq++;
							g_run_end = qq;
							continue;
						}
						if (!g_nonempty) {  // next batch of descriptors
							if (g_base >= nn) return false;
							d_n = 0;
							if (g_base + lane < nn) {
								const int64_t cc = (int64_t)nlist[g_base + lane] * a.nz + s;
								const int64_t c0 = a.cell_start[cc], c1 = a.cell_start[cc + 1];
								const CellInfo ci = a.cinfo[cc];
								d_start = c0;
								d_n = (int)(c1 - c0);
								d_label = ci.label;
								d_nlab = ci.nlab;
								d_umin = ci.umin;
								d_umax = ci.umax;
								d_vmin = ci.vmin;
								d_vmax = ci.vmax;
							}
							g_nonempty = __ballot_sync(0xffffffffu, d_n > 0);
							g_base += 32;
							continue;
						}
						const int e = __ffs(g_nonempty) - 1;
						g_nonempty &= g_nonempty - 1u;
						const double umin = __shfl_sync(0xffffffffu, d_umin, e), umax = __shfl_sync(0xffffffffu, d_umax, e);
						const double vmin = __shfl_sync(0xffffffffu, d_vmin, e), vmax = __shfl_sync(0xffffffffu, d_vmax, e);
						// per-warp culling against the cell's bounding box: can any lane have a pair in this window?
						double ulo = __dsub_rn(p.u, umax), uhi = __dsub_rn(p.u, umin);
						double vlo = __dsub_rn(p.v, vmax), vhi = __dsub_rn(p.v, vmin);
						bool nocull = false;
						double su = 0.0, sv = 0.0;
						if (periodic) {
							if (!(ulo >= -halfL && uhi <= halfL)) {
								if (ulo > halfL) {  // every pair of this lane with the cell wraps down
									ulo = __dsub_rn(ulo, L);
									uhi = __dsub_rn(uhi, L);
									su = -L;
								} else if (uhi < -halfL) {
									ulo = __dadd_rn(ulo, L);
									uhi = __dadd_rn(uhi, L);
									su = L;
								} else {
									nocull = true;  // the cell straddles +-L/2 for this lane (tiny boxes only)
								}
							}
							if (!(vlo >= -halfL && vhi <= halfL)) {
								if (vlo > halfL) {
									vlo = __dsub_rn(vlo, L);
									vhi = __dsub_rn(vhi, L);
									sv = -L;
								} else if (vhi < -halfL) {
									vlo = __dadd_rn(vlo, L);
									vhi = __dadd_rn(vhi, L);
									sv = L;
								} else {
									nocull = true;
								}
							}
						}
						const double mu = ulo > 0.0 ? ulo : (uhi < 0.0 ? -uhi : 0.0);
						const double mv = vlo > 0.0 ? vlo : (vhi < 0.0 ? -vhi : 0.0);
						const double dmin2 = __dadd_rn(__dmul_rn(mu, mu), __dmul_rn(mv, mv));
						const bool need = !zw.dead && (nocull || dmin2 < win_hi);
						if (!__any_sync(0xffffffffu, need)) continue;  // never even staged
						g_xy = __any_sync(0xffffffffu, !zw.dead && nocull) ? 2
							   : (__any_sync(0xffffffffu, !zw.dead && (su != 0.0 || sv != 0.0)) ? 1 : 0);
						g_su = su;
						g_sv = sv;
						g_pos = __shfl_sync(0xffffffffu, d_start, e);
						g_cell_end = g_pos + __shfl_sync(0xffffffffu, d_n, e);
						g_label = __shfl_sync(0xffffffffu, d_label, e);
						g_nlab = __shfl_sync(0xffffffffu, d_nlab, e);
						g_run_end = g_cell_end;
						if (g_nlab > 1) {
							g_label = a.cand_jk[g_pos];
							long long qq = g_pos + 1;
							while (qq < g_cell_end && a.cand_jk[qq] == g_label) qq++;
							g_run_end = qq;
						}
					}
				};

				// ---- software pipeline: issue the bulk copy of chunk k+1, then work on chunk k ----------------------------
				Chunk pend, nxt;
				double pend_su = 0.0, pend_sv = 0.0, nxt_su = 0.0, nxt_sv = 0.0;  // per-lane image shifts of the chunk's cell
				pend.n = 0;
				pend.label = -1;
				int pend_st = 0, cur_label = -1;
				bool more = true;
				while (more || pend.n > 0) {
					bool got = false;
					if (more) {
						got = next_chunk(nxt);
						more = got;
						nxt_su = g_su;
						nxt_sv = g_sv;
					}
					if (got && lane == 0) {
						const uint32_t bytes = (uint32_t)nxt.n * (uint32_t)sizeof(Cand);
						mbar_expect_tx(&my_full[st_issue], bytes);
						bulk_load(my_ring + (size_t)st_issue * CH, a.cand + nxt.start, bytes, &my_full[st_issue]);
					}
					// a label change (or the end of the window: pend.n == 0) flushes the private slots
					const int lab = (pend.n > 0) ? pend.label : -2;
					if (lab != cur_label) {
						if (cur_label >= 0) flush(cur_label);
						cur_label = lab;
					}
					if (pend.n > 0) {
						if (pend_st == 0) {
							mbar_wait(&my_full[0], phase0);
							phase0 ^= 1u;
						} else {
							mbar_wait(&my_full[1], phase1);
							phase1 ^= 1u;
						}
						if (!zw.dead) tested += (unsigned long long)pend.n;
						const uint32_t cb = my_ring_u32 + (uint32_t)pend_st * (uint32_t)(CH * sizeof(Cand));
						bool susp;
						if (pend.xy == 2)  // tiny boxes: a column pair straddles +-L/2
							susp = pair_loop<UNITW, false, false, true>(cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1, 0.0,
																		 0.0, rw, hi_lane, zw, acc);
						else if (pend.xy == 1 && warp_zg)
							susp = pair_loop<UNITW, true, true, false>(cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1,
																		pend_su, pend_sv, rw, hi_lane, zw, acc);
						else if (pend.xy == 1)
							susp = pair_loop<UNITW, true, false, false>(cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1,
																		 pend_su, pend_sv, rw, hi_lane, zw, acc);
						else if (warp_zg)
							susp = pair_loop<UNITW, false, true, false>(cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1, 0.0,
																		 0.0, rw, hi_lane, zw, acc);
						else  // the common case: no periodic image in the projected axes, constant one along the line of sight
							susp = pair_loop<UNITW, false, false, false>(cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1,
																		  0.0, 0.0, rw, hi_lane, zw, acc);
						if (__any_sync(0xffffffffu, susp))
							slow_pairs<UNITW>(susp, cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1, rw, hi_lane, zw.t_lo,
											  zw.t_hi, zw.t_split, acc, nan_pairs);
						__syncwarp();  // every lane is done with the stage before it is refilled
					}
					if (got) {
						pend = nxt;
						pend_su = nxt_su;
						pend_sv = nxt_sv;
						pend_st = st_issue;
						st_issue ^= 1;
					} else {
						pend.n = 0;
					}
				}
				if (cur_label >= 0) flush(cur_label);  // (normally already flushed by the drain iteration)
			}
		}
	}

	// ---- statistics ------------------------------------------------------------------------------------------------------
	for (int o = 16; o > 0; o >>= 1) {
		tested += __shfl_down_sync(0xffffffffu, tested, o);
		binned += __shfl_down_sync(0xffffffffu, binned, o);
		nan_pairs += __shfl_down_sync(0xffffffffu, nan_pairs, o);
	}
	if (lane == 0) {
		atomicAdd(&a.A.stats[0], tested);
		atomicAdd(&a.A.stats[1], binned);
		atomicAdd(&a.A.stats[2], nan_pairs);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------------------
inline int tiled_prepare_candidates(const TiledConfig &cfg, const GridDims &g, const DevParams &P, const uint32_t *,
									const Cand *cand, const int32_t *cand_jk, int64_t, const int64_t *cell_start,
									void *ws, cudaStream_t st) {
	TiledWorkspace w = carve_tiled(cfg, g, ws);
	MIA_CUDA_CHECK(cudaMemsetAsync(w.slab_lo, 0xFF, sizeof(double) * cfg.nz, st));
	MIA_CUDA_CHECK(cudaMemsetAsync(w.slab_hi, 0x00, sizeof(double) * cfg.nz, st));
	const int64_t ncell = g.ncell();
	k_cell_info<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>((const unsigned char *)cand, cfg.sym ? cfg.sym : (int)sizeof(Cand), cand_jk,
																 cell_start, ncell, cfg.nz, w.cinfo, (unsigned long long *)w.slab_lo,
																 (unsigned long long *)w.slab_hi, g.order, g.ncv);
	MIA_CUDA_CHECK(cudaGetLastError());
	if (cfg.v2) {
		const int rc = rppi2_prepare(cfg, g, w, st);
		if (rc) return rc;
	} else if (cfg.geom == MIA_GEOM_RMU) {
		const int64_t ncol = (int64_t)g.ncu * g.ncv;
		k_col_info<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(w.cinfo, ncol, cfg.nz, cfg.n_lr, w.colinfo, w.colreg,
																   w.slab_lo, w.slab_hi);
		MIA_CUDA_CHECK(cudaGetLastError());
	} else {
		MIA_CUDA_CHECK(cudaMemcpyAsync(w.lut, cfg.lut, sizeof(LutEntry) * LUT_SIZE, cudaMemcpyHostToDevice, st));
	}
	(void)P;
	return 0;
}

inline int tiled_launch(const TiledConfig &cfg, const GridDims &gc, const GridDims &g, const DevParams &P, const Grid &G, const Prim *prim,
						const int64_t *prim_cell_start, int64_t nS, bool unit_w, mia_shard shard, const Accum &A, void *ws,
						int *flags, cudaStream_t st, cudaEvent_t ev_before = nullptr, cudaEvent_t ev_after = nullptr) {
	TiledWorkspace w = carve_tiled(cfg, gc, ws);  // carved on the candidate grid (gc); g = grid of the shape sort
	const int64_t ncol = (int64_t)g.ncu * g.ncv;
	if (nS == 0 || G.n_cand == 0) {
		if (ev_before) MIA_CUDA_CHECK(cudaEventRecord(ev_before, st));
		if (ev_after) MIA_CUDA_CHECK(cudaEventRecord(ev_after, st));
		return 0;
	}
	// ---- task table -------------------------------------------------------------------------------------------------------
	MIA_CUDA_CHECK(cudaMemsetAsync(w.task_cost, 0, sizeof(unsigned long long) * cfg.max_tasks, st));
	const int nzs = cfg.nz * (g.sub > 1 ? g.sub * g.sub : 1);
	// when there are few tasks per worker warp (small catalogues, many GPUs) cut each task along the line of sight
	const int spt = 32 / cfg.hsplit;  // shape galaxies per warp task
	const double tasks_est = (double)nS / (double)spt + 0.5 * (double)ncol;
	// (measured, profiles/r01_tuning.md: finer tasks help the (r_p, Pi) kernel's tail, the (r, mu_r) kernel pays more per task)
	const char *tpw_env = getenv("MIA_TASKS_PER_WARP");
	// (cell-by-cell kernel: static slots, finer tasks shorten its tail; the others hand out slots dynamically and reuse per-task
	// set-up across slabs, so they prefer long tasks: cfg2 96.7 ms at 8 against 98.3 at 32)
	const double tasks_per_warp = tpw_env ? atof(tpw_env) : ((cfg.geom == MIA_GEOM_RMU || cfg.v2) ? 8.0 : MIA_TASKS_PER_WARP);
	const double want = tasks_per_warp * (double)cfg.n_ctas * TW * (double)shard.count;
	int split = (int)ceil(want / (tasks_est > 1.0 ? tasks_est : 1.0));
	split = split < 1 ? 1 : (split > MAX_SPLIT ? MAX_SPLIT : split);
	if (cfg.geom != MIA_GEOM_RMU && split > cfg.nz) split = cfg.nz;  // (r_p, Pi): parts = ranges of slabs; (r, mu_r): of columns
	k_col_chunks<<<(unsigned)((ncol + 1 + 255) / 256), 256, 0, st>>>(prim_cell_start, ncol, nzs, split, w.col_chunks, spt);
	MIA_CUDA_CHECK(cudaGetLastError());
	size_t cb = w.cub_bytes;
	MIA_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.col_chunks, w.task_off, (int)(ncol + 1), st));
	TiledArgs a;
	a.P = P;
	a.ratio = cfg.ratio;
	a.hsplit = cfg.hsplit;
	if (cfg.v2) {
		const int rc = rppi2_fill_tasks(a, prim_cell_start, G.cell_start, w.task_off, (int)ncol, nzs, P.ku, split, cfg.sym, w.task_col,
										w.task_first, w.task_n, w.task_slab, w.task_cost, w.n_tasks, st);
		if (rc) return rc;
	} else if (cfg.geom == MIA_GEOM_RMU) {
		const int rc = rmu_fill_tasks(a, prim_cell_start, G.cell_start, w.task_off, (int)ncol, nzs, P.ku, split, cfg.sym, w.task_col, w.task_first,
									  w.task_n, w.task_slab, w.task_cost, w.n_tasks, st);
		if (rc) return rc;
	} else {
		const double cs = P.L / g.ncu, reach = sqrt(P.r2_thr[P.n_r]) * (1.0 + 1e-6);
		k_fill_tasks<<<(unsigned)((ncol + 255) / 256), 256, 0, st>>>(prim_cell_start, G.cell_start, w.task_off, g.ncu, g.ncv,
																	 cfg.nz, nzs, split, P.ku, P.periodic, cs, reach, w.task_col,
																	 w.task_first, w.task_n, w.task_slab, w.task_cost, w.n_tasks);
	}
	MIA_CUDA_CHECK(cudaGetLastError());
	cb = w.cub_bytes;
	MIA_CUDA_CHECK(cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.task_cost, w.task_cum, cfg.max_tasks, st));

	// ---- launch.  The number of worker warps is fixed (SLOTS_PER_SM CTAs per SM), NOT derived from occupancy: the
	// grouping of the fp64 sums, hence every output bit, is the same for the weighted and the unit-weight kernel variants,
	// which is what lets w = 0.5 scale the results by exactly 1/4 (reference tests/test_weights.py:34-35). ---------------
	const bool rmu = cfg.geom == MIA_GEOM_RMU;
	const bool sig = cfg.sig != 0;
	if (sig && !rmu && !cfg.v2) return MIA_ERR_UNSUPPORTED;  // (plan_tiled never selects the cell-by-cell kernel for such calls)
	a.ch_sym = (rmu && cfg.sym) ? rmu_sym_chunk(unit_w, cfg.w_r * P.n_2) : 0;
	const size_t smem = rmu ? (cfg.sym ? tiled_rmu_sym_smem_bytes(unit_w, cfg.w_r * P.n_2, a.ch_sym) : tiled_rmu_smem_bytes(unit_w, sig))
							: (cfg.sym ? tiled_rppi2s_smem_bytes(unit_w)
									   : (cfg.v2 ? tiled_rppi2_smem_bytes(unit_w, sig) : tiled_smem_bytes(unit_w)));
	if (rmu || cfg.v2) {
	} else if (unit_w) {
		MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rppi<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	} else {
		MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rppi<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	}
	a.colinfo = w.colinfo;
	a.colreg = w.colreg;
	a.vlo = w.vlo;
	a.vhi = w.vhi;
	a.n_lr = cfg.n_lr;
	a.w_r = cfg.w_r;
	a.cand = G.cand;
	a.cand_jk = G.cand_jk;
	a.cell_start = G.cell_start;
	a.cinfo = w.cinfo;
	a.slab_lo = w.slab_lo;
	a.slab_hi = w.slab_hi;
	a.prim = prim;
	a.task_col = w.task_col;
	a.task_first = w.task_first;
	a.task_n = w.task_n;
	a.task_slab = w.task_slab;
	a.task_cost = w.task_cost;
	a.task_cum = w.task_cum;
	a.n_tasks = w.n_tasks;
	a.lut = w.lut;
	a.lut_hi0 = cfg.lut_hi0;
	a.lut_shift = cfg.lut_shift;
	a.lut_n = cfg.lut_n;
	a.A = A;
	a.nz = cfg.nz;
	a.n_side = cfg.n_side;
	a.n_workers = cfg.n_partials;  // worker slots = accumulator copies
	a.shard_index = shard.index;
	a.shard_count = shard.count;
	a.max_tasks = cfg.max_tasks;
	a.flags = flags;
	if (ev_before) MIA_CUDA_CHECK(cudaEventRecord(ev_before, st));
	if (rmu) {
		const int rc = cfg.sym ? launch_rmu_sym(a, unit_w, P.los == 2, cfg.n_ctas, smem, st)
							   : launch_rmu(a, unit_w, P.los == 2, sig, cfg.n_ctas, smem, st);
		if (rc) return rc;
	} else if (cfg.sym) {
		const int rc = launch_rppi2s(a, unit_w, cfg.n_ctas, smem, st);
		if (rc) return rc;
	} else if (cfg.v2) {
		const int rc = launch_rppi2(a, unit_w, sig, cfg.n_ctas, smem, st);
		if (rc) return rc;
	} else if (unit_w) {
		k_tiled_rppi<true><<<cfg.n_ctas, TP, smem, st>>>(a);
	} else {
		k_tiled_rppi<false><<<cfg.n_ctas, TP, smem, st>>>(a);
	}
	MIA_CUDA_CHECK(cudaGetLastError());
	if (ev_after) MIA_CUDA_CHECK(cudaEventRecord(ev_after, st));
	return 0;
}

}  // namespace mia

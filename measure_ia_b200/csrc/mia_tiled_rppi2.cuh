// mia_tiled_rppi2.cuh -- ROW-STREAMING variant of the tiled (r_p, Pi) pair kernel for sm_100a.
//
// Same pair loop, private windowed histograms, z-window logic and fixed-order reductions as mia_tiled.cuh (it reuses
// pair_loop / slow_pairs / flush_slots / z_window from there; reference: src/measureia/measure_w_box_jk.py:387-461).
// What changes is how candidates reach the loop -- the design that made the (r, mu_r) kernel fast:
//   * candidates are sorted by (u row, slab, v cell), so for one slab the cells of a u row are CONTIGUOUS along v.  For every
//     u row within reach the warp streams ONE contiguous range: the v cells within sqrt(r_hi^2 - d_u^2) of its shapes (exact
//     union over the shapes that are alive for the slab), in chunks of <= CH, instead of one small chunk per cell;
//   * the cells are finer (r_max / 8 instead of r_max / 4) and the shape columns 2 x 2 cells wide: culling in both
//     projected axes at half the old granularity at no cost in chunk size;
//   * up to 32 rows are opened lane-parallel in a ROUND (range, candidate offsets from cell_start, label from the per-(row,
//     v region) table, warp-constant periodic-image codes from bounding boxes) and consumed by a __noinline__ function
//     with its own register allocation.
// The grid is aligned with the jackknife sub-boxes (a cell carries one label); rows whose region holds several labels
// (unaligned grids) take a cell-by-cell path.
#pragma once
#include "mia_tiled.cuh"
#include "mia_tiled_rmu.cuh"

namespace mia {

#ifndef MIA_RPPI2_DIV
#define MIA_RPPI2_DIV 10
#endif
#ifndef MIA_RPPI2_RATIO
#define MIA_RPPI2_RATIO 2
#endif

inline size_t tiled_rppi2_smem_bytes(bool unit_w, bool sig) {
	const size_t fixed = sizeof(Cand) * TW * STAGES * CH + 256 + 768;
	const size_t per_slot = (size_t)TP * (8 + 8 + 4 + (unit_w ? 0 : 8) + (sig ? 8 : 0));
	return fixed + per_slot * NSLOT;
}

// Columns of about r_max / DIV, a multiple of lcm(n_side, ratio) per side.
inline bool plan_rppi2_grid(const mia_params *p, int n_side, TiledConfig &cfg, int &nc, int nz, int &k) {
	const double L = p->boxsize, reach = p->r_search * (1.0 + 1e-6);
	int div = env_int("MIA_RPPI2_DIV", cfg.w_r > 0 ? cfg.w_r : MIA_RPPI2_DIV), ratio = env_int("MIA_RPPI2_RATIO", MIA_RPPI2_RATIO);
	cfg.w_r = 0;
	if (div < 1) div = 1;
	if (ratio < 1) ratio = 1;
	for (;; div--) {
		nc = (int)floor(L / (reach / (double)div));
		if (nc > 2048) nc = 2048;
		if (nc < 1) nc = 1;
		int rt = ratio;
		while (rt > 1 && nc < 4 * rt) rt--;
		int m = rt;
		if (n_side > 1) {
			int a = n_side, b = rt;
			while (b) {
				const int t_ = a % b;
				a = b;
				b = t_;
			}
			m = n_side / a * rt;
		}
		if (nc >= 2 * m) nc = nc / m * m;
		else if (nc >= 2 * rt) nc = nc / rt * rt;
		else rt = 1;
		const double cs = L / nc;
		k = (int)ceil(reach / cs);
		if (k < 1) k = 1;
		const unsigned long long nkeys =
			(unsigned long long)nc * nc * nz * 4ull * (unsigned long long)(p->num_jk > 0 ? p->num_jk : 1);
		if (nkeys <= (1ull << 31)) {
			cfg.ratio = rt;
			break;
		}
		if (div == 1) return false;
	}
	cfg.n_lr = (n_side > 1 && nc % n_side == 0) ? n_side : 1;  // regions along v
	cfg.v2 = 1;
	return true;
}

// Per (u row, slab): bounding box in u, and per v region the label its candidates share (-1 several, -2 none).  Also
// accumulates the v bounds per column index cv (turned into envelopes by k_v_envelope).
__global__ void k_row_info(const CellInfo *__restrict__ cinfo, int64_t nrow, int ncv, int n_vr, ColInfo *__restrict__ out,
						   int32_t *__restrict__ rowreg, unsigned long long *__restrict__ vlo, unsigned long long *__restrict__ vhi) {
	const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (r >= nrow) return;
	ColInfo o;
	o.umin = o.vmin = INFINITY;
	o.umax = o.vmax = -INFINITY;
	const int per = ncv / n_vr;
	for (int g = 0; g < n_vr; g++) {
		int lab = -2;
		const int c1 = (g == n_vr - 1) ? ncv : (g + 1) * per;
		for (int cv = g * per; cv < c1; cv++) {
			const CellInfo ci = cinfo[r * ncv + cv];
			if (ci.nlab == 0) continue;
			o.umin = fmin(o.umin, ci.umin);
			o.umax = fmax(o.umax, ci.umax);
			o.vmin = fmin(o.vmin, ci.vmin);
			o.vmax = fmax(o.vmax, ci.vmax);
			atomicMin(&vlo[cv], (unsigned long long)__double_as_longlong(ci.vmin + 0.0));  // v >= 0: bit pattern monotone
			atomicMax(&vhi[cv], (unsigned long long)__double_as_longlong(ci.vmax + 0.0));
			if (ci.nlab > 1) lab = -1;
			else if (lab == -2) lab = ci.label;
			else if (lab != ci.label) lab = -1;
		}
		rowreg[r * n_vr + g] = lab;
	}
	out[r] = o;
}

// vlo[c] = smallest v in columns >= c, vhi[c] = largest v in columns <= c: valid bounds for any run of columns.
__global__ void k_v_envelope(double *__restrict__ vlo, double *__restrict__ vhi, int ncv) {
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	double cur = INFINITY;
	for (int c = ncv - 1; c >= 0; c--) {
		const double v = vlo[c];
		if (v == v) cur = fmin(cur, v);  // untouched entries hold the NaN fill pattern
		vlo[c] = cur;
	}
	cur = -INFINITY;
	for (int c = 0; c < ncv; c++) {
		cur = fmax(cur, vhi[c]);  // untouched entries hold 0.0 <= every coordinate
		vhi[c] = cur;
	}
}

// One warp task = up to 32 consecutive shape galaxies of one shape column (x a range of slabs when tasks are scarce);
// cost = shapes x candidates in the u rows within reach.
__global__ void k_fill_tasks_rppi2(const int64_t *__restrict__ prim_cell_start, const int64_t *__restrict__ cell_start,
								   const int32_t *__restrict__ task_off, int ncu, int ncv, int nz, int ratio, int nzs, int split,
								   int k, int periodic, int sym, int32_t *__restrict__ task_col, int64_t *__restrict__ task_first,
								   int32_t *__restrict__ task_n, int32_t *__restrict__ task_slab,
								   unsigned long long *__restrict__ task_cost, int32_t *__restrict__ n_tasks) {
	const int ncu_s = ncu / ratio, ncv_s = ncv / ratio;
	const int64_t ncol = (int64_t)ncu_s * ncv_s;
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c == 0) n_tasks[0] = task_off[ncol];
	if (c >= ncol) return;
	const int64_t p0 = prim_cell_start[c * nzs], p1 = prim_cell_start[(c + 1) * nzs];
	if (p1 <= p0) return;
	const int su0 = (int)(c / ncv_s);
	const bool all_u = 2 * k + ratio >= ncu;
	const int nrows = all_u ? ncu : 2 * k + ratio;
	unsigned long long W = 0;
	for (int o = sym ? k : 0; o < nrows; o++) {  // symmetric kernel: own rows and the rows ahead only
		int cu = all_u ? o : ratio * su0 - k + o;
		if (cu < 0 || cu >= ncu) {
			if (!periodic) continue;
			cu = (cu + ncu) % ncu;
		}
		W += (unsigned long long)(cell_start[(int64_t)(cu + 1) * nz * ncv] - cell_start[(int64_t)cu * nz * ncv]);
	}
	W = W * (unsigned long long)(2 * k + ratio) / (unsigned long long)(ncv > 0 ? ncv : 1) + 1ull;  // share of v within reach
	int t = task_off[c];
	for (int64_t p = p0; p < p1; p += 32) {
		const int n = (int)((p1 - p < 32) ? (p1 - p) : 32);
		for (int part = 0; part < split; part++, t++) {
			task_col[t] = (int32_t)c;
			task_first[t] = p;
			task_n[t] = n;
			task_slab[2 * t] = (int)((long long)nz * part / split);
			task_slab[2 * t + 1] = (int)((long long)nz * (part + 1) / split);
			// the loop is branch-free over the 32 lanes: a task with few shapes costs almost what a full one does (only its
			// streamed ranges shrink a little), so the cost is mostly the candidates in reach, not shapes x candidates
			task_cost[t] = (unsigned long long)(16 + n / 2) * W / (unsigned long long)split + 1ull;
		}
	}
}

// Everything the chunk consumer needs for one (task, slab, window); lives in the kernel's local memory.
struct R2Ctx {
	// per lane
	double pu, pv, pl, a0, a1, hi_lane, pe, pw;
	ZWindow zw;
	unsigned key;
	// per (task, slab, window)
	double L, halfL;
	RWindow rw;
	int rb, periodic, warp_zg;
	PrivAcc acc;
	uint32_t ring_u32;
	const Cand *cand;
	Cand *ring;
	uint64_t *full;
	FlushCtx fc;
	// mutable
	uint32_t phase0, phase1;
	int st_issue, cur_label;
	unsigned long long tested, binned, nan_pairs;
};

// Consume one round: lane e (bit e of mask) holds a row descriptor = up to two contiguous candidate ranges with ONE
// jackknife label and warp-constant image codes (codes = cu | cv(piece A) << 2 | cv(piece B) << 4; 0 none, 1: d -= L,
// 2: d += L, 3: straddles +-L/2 -> wrap per pair).
#ifndef MIA_RPPI2_INLINE
#define MIA_RPPI2_INLINE 1
#endif
#if MIA_RPPI2_INLINE
#define MIA_R2_ATTR __forceinline__
#else
#define MIA_R2_ATTR __noinline__
#endif
template <bool UNITW, bool SIG>
__device__ MIA_R2_ATTR void process_round2(R2Ctx *cx, int sA, int eA, int sB, int eB, int lab, int codes, unsigned mask) {
	const int lane = threadIdx.x & 31;
	const double L = cx->L, halfL = cx->halfL, pu = cx->pu, pv = cx->pv, pl = cx->pl, a0 = cx->a0, a1 = cx->a1;
	const double hi_lane = cx->hi_lane;
	const RWindow rw = cx->rw;
	const ZWindow zw = cx->zw;
	const PrivAcc acc = cx->acc;
	const int periodic = cx->periodic;
	const bool warp_zg = cx->warp_zg != 0;
	const uint32_t ring_u32 = cx->ring_u32;
	const Cand *cand = cx->cand;
	Cand *ring = cx->ring;
	uint64_t *full = cx->full;
	uint32_t phase0 = cx->phase0, phase1 = cx->phase1;
	int st_issue = cx->st_issue, cur_label = cx->cur_label;
	unsigned tested = 0, binned = 0;  // per round: < 2^32

	int pend_n = 0, pend_st = 0, pend_label = -1, pend_codes = 0;
	auto consume = [&]() {
		if (pend_label != cur_label) {
			if (cur_label >= 0)
				binned += flush_slots<UNITW, SIG>(cx->fc, acc, cx->key, zw.dead, cx->pe, cx->pw, rw.ra, cx->rb, cur_label);
			cur_label = pend_label;
		}
		const int cu_ = pend_codes & 3, cv_ = (pend_codes >> 2) & 3;
		if (pend_st == 0) {
			mbar_wait(&full[0], phase0);
			phase0 ^= 1u;
		} else {
			mbar_wait(&full[1], phase1);
			phase1 ^= 1u;
		}
		if (!zw.dead) tested += (unsigned)pend_n;
		const uint32_t cb = ring_u32 + (uint32_t)pend_st * (uint32_t)(CH * sizeof(Cand));
		const double su = code_shift(cu_, L), sv = code_shift(cv_, L);
		bool susp;
		if (cu_ == 3 || cv_ == 3)
			susp = pair_loop<UNITW, false, false, true, SIG>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, rw, hi_lane, zw, acc);
		else if ((cu_ | cv_) && warp_zg)
			susp = pair_loop<UNITW, true, true, false, SIG>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, su, sv, rw, hi_lane, zw, acc);
		else if (cu_ | cv_)
			susp = pair_loop<UNITW, true, false, false, SIG>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, su, sv, rw, hi_lane, zw, acc);
		else if (warp_zg)
			susp = pair_loop<UNITW, false, true, false, SIG>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, rw, hi_lane, zw, acc);
		else
			susp = pair_loop<UNITW, false, false, false, SIG>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, rw, hi_lane, zw, acc);
		if (__any_sync(0xffffffffu, susp))
			slow_pairs<UNITW, SIG>(susp, cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, rw, hi_lane, zw.t_lo, zw.t_hi, zw.t_split,
								   acc, cx->nan_pairs);
		__syncwarp();
	};

	while (mask) {
		const int e = __ffs(mask) - 1;
		mask &= mask - 1u;
		const int d_lab = __shfl_sync(0xffffffffu, lab, e), d_codes = __shfl_sync(0xffffffffu, codes, e);
#pragma unroll 1
		for (int piece = 0; piece < 2; piece++) {
			int s = __shfl_sync(0xffffffffu, piece ? sB : sA, e);
			const int en = __shfl_sync(0xffffffffu, piece ? eB : eA, e);
			const int pc = (d_codes & 3) | (((d_codes >> (2 + 2 * piece)) & 3) << 2);
			while (s < en) {
				const int rest = en - s, nch = (rest + CH - 1) / CH;
				const int n = (rest + nch - 1) / nch;
				if (lane == 0) {
					const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(Cand);
					mbar_expect_tx(&full[st_issue], bytes);
					bulk_load(ring + (size_t)st_issue * CH, cand + s, bytes, &full[st_issue]);
				}
				if (pend_n > 0) consume();
				pend_n = n;
				pend_st = st_issue;
				pend_label = d_lab;
				pend_codes = pc;
				st_issue ^= 1;
				s += n;
			}
		}
	}
	if (pend_n > 0) consume();
	cx->phase0 = phase0;
	cx->phase1 = phase1;
	cx->st_issue = st_issue;
	cx->cur_label = cur_label;
	cx->tested += tested;
	cx->binned += binned;
}

// Out-of-line copy for the rare cell-by-cell path: the consumer is inlined ONCE into the kernel (three inlined copies of
// its five loop variants cost 0.6 stall cycles per instruction in instruction-cache misses).
template <bool UNITW, bool SIG>
__device__ __noinline__ void process_round2_cold(R2Ctx *cx, int sA, int eA, int sB, int eB, int lab, int codes, unsigned mask) {
	process_round2<UNITW, SIG>(cx, sA, eA, sB, eB, lab, codes, mask);
}

// SIG: also accumulate sum (w_D w_S e+)^2 per bin (the `variance` of the reference's brute variants, measure_w_box_jk.py:196)
template <bool UNITW, bool SIG>
__global__ void __launch_bounds__(TP, 3) k_tiled_rppi2(const TiledArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	const DevParams &P = a.P;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb = P.n_r * P.n_2;
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	const int periodic = P.periodic;
	const double L = P.L, halfL = P.halfL;
	const int nz = a.nz, ncu = P.ncu, ncv = P.ncv, ratio = a.ratio, kk = P.ku;

	// ---- shared memory carve-up ------------------------------------------------------------------------------------
	Cand *ring = reinterpret_cast<Cand *>(smem);  // [warp][stage][CH]
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + sizeof(Cand) * TW * STAGES * CH);  // [warp][stage]
	double *thr2_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(full) + 256);
	unsigned char *accbase = reinterpret_cast<unsigned char *>(thr2_s) + 768;
	const uint32_t acc_u32 = smem_u32(accbase);
	Cand *my_ring = ring + (size_t)warp * STAGES * CH;

	R2Ctx cx;
	cx.acc.a2 = acc_u32 + (uint32_t)tid * 16u;
	cx.acc.aw = acc_u32 + (uint32_t)NSLOT * TP * 16u + (uint32_t)tid * 8u;
	cx.acc.av = acc_u32 + (uint32_t)NSLOT * TP * (UNITW ? 16u : 24u) + (uint32_t)tid * 8u;
	cx.acc.ac = acc_u32 + (uint32_t)NSLOT * TP * ((UNITW ? 16u : 24u) + (SIG ? 8u : 0u)) + (uint32_t)tid * 4u;
	cx.ring_u32 = smem_u32(my_ring);
	cx.ring = my_ring;
	cx.full = full + warp * STAGES;
	cx.cand = a.cand;
	cx.L = L;
	cx.halfL = halfL;
	cx.periodic = periodic;
	cx.phase0 = cx.phase1 = 0u;
	cx.st_issue = 0;
	cx.cur_label = -1;
	cx.tested = cx.binned = cx.nan_pairs = 0ull;

	if (tid == 0) {
		for (int s = 0; s < TW * STAGES; s++) mbar_init(&full[s], 1);
		mbar_fence_init();
		if (blockIdx.x == 0) a.A.stats[6] = (unsigned long long)a.n_tasks[0];
	}
	for (int e = tid; e <= P.n_2; e += blockDim.x) thr2_s[e] = P.thr2[e];
#pragma unroll
	for (int s = 0; s < NSLOT; s++) {
		sts_v2(cx.acc.a2 + (uint32_t)s * TP * 16u, 0.0, 0.0);
		if (!UNITW) sts_f64(cx.acc.aw + (uint32_t)s * TP * 8u, 0.0);
		if (SIG) sts_f64(cx.acc.av + (uint32_t)s * TP * 8u, 0.0);
		sts_u32(cx.acc.ac + (uint32_t)s * TP * 4u, 0u);
	}
	__syncthreads();  // the only CTA-wide synchronisation

	// ---- this warp's share of the tasks ----------------------------------------------------------------------------------
	for (int slot = next_slot(a.flags + 2); slot < a.n_workers; slot = next_slot(a.flags + 2)) {  // (body not re-indented)
	int task0 = 0, task1 = 0;
	{
		const int nt = a.n_tasks[0];
		if (nt > 0) {
			const double total2 = 2.0 * (double)a.task_cum[nt - 1];
			const int RG = a.shard_count * a.n_workers;
			const int mine = slot * a.shard_count + a.shard_index;  // slots interleaved across ranks
			auto slot_of = [&](int t) {
				const double mid2 = 2.0 * (double)a.task_cum[t] - (double)a.task_cost[t];
				int s = (int)(mid2 / total2 * (double)RG);
				return s < RG - 1 ? s : RG - 1;
			};
			auto lower = [&](int target) {
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

	const size_t part = (size_t)slot * (size_t)a.A.rows * nb;  // the slot's accumulator copy
	cx.fc.pcnt = a.A.cnt + part;
	cx.fc.pddw = a.A.ddw + part;
	cx.fc.psp = a.A.sp + part;
	cx.fc.psc = a.A.sc + part;
	cx.fc.pvar = SIG ? a.A.var + (size_t)slot * nb : nullptr;
	cx.fc.flags = a.flags;
	cx.fc.n_2 = P.n_2;
	cx.fc.nb = nb;
	cx.fc.J = J;
	cx.fc.num_jk = P.num_jk;

	ZParams zp;
	zp.thr2 = thr2_s;
	zp.L = L;
	zp.halfL = halfL;
	zp.n_2 = P.n_2;
	zp.periodic = periodic;

	const int n_win = (P.n_r + W_R - 1) / W_R;
	const int n_vr = a.n_lr, vr_cells = ncv / n_vr;
	const int ncv_s = ncv / ratio;
	const bool all_u = 2 * kk + ratio >= ncu;
	const int nrows = all_u ? ncu : 2 * kk + ratio;
	const double eps_v = 1e-9 * L;

	auto axis_code = [&](double b0, double b1, double cmin, double cmax) -> int {
		if (!periodic) return 0;
		const double lo = __dsub_rn(b0, cmax), hi = __dsub_rn(b1, cmin);
		if (lo >= -halfL && hi <= halfL) return 0;
		if (lo > halfL) return 1;   // every pair wraps down: sep -= L (measure_w_box_jk.py:403)
		if (hi < -halfL) return 2;  // every pair wraps up (:404)
		return 3;
	};
	auto gap = [&](double x, double cmin, double cmax, int code) -> double {
		if (code == 3) {
			double g = fmax(0.0, fmax(cmin - x, x - cmax));
			g = fmin(g, fmax(0.0, fmax((cmin + L) - x, x - (cmax + L))));
			return fmin(g, fmax(0.0, fmax((cmin - L) - x, x - (cmax - L))));
		}
		const double sh = code == 1 ? L : (code == 2 ? -L : 0.0);
		return fmax(0.0, fmax((cmin + sh) - x, x - (cmax + sh)));
	};

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
		cx.pu = p.u;
		cx.pv = p.v;
		cx.pl = p.l;
		cx.a0 = p.a0;
		cx.a1 = p.a1;
		cx.pe = p.w * p.e;
		cx.pw = p.w;
		const int su0 = col / ncv_s;

		const int slab0 = a.task_slab[2 * task], slab1 = a.task_slab[2 * task + 1];
		// per-row v ranges of the two top windows, valid while the same lanes are alive (see the round set-up below)
		float c_vmin0 = 0.f, c_vmax0 = 0.f, c_vmin1 = 0.f, c_vmax1 = 0.f;
		int c_code0 = 0, c_code1 = 0;
		unsigned c_alive0 = 0u, c_alive1 = 0u;
		for (int s = slab0; s < slab1; s++) {
			const double zlo = a.slab_lo[s], zhi = a.slab_hi[s];
			if (!(zlo <= zhi)) continue;  // empty slab (uniform branch)
			ZWindow zw = z_window(p.l, zlo, zhi, zp);
			if (!active) zw.dead = true;
			if (active && zw.err) atomicExch(&a.flags[1], 1);
			const unsigned alive = __ballot_sync(0xffffffffu, !zw.dead);
			if (!alive) continue;
			cx.zw = zw;
			cx.warp_zg = __any_sync(0xffffffffu, !zw.dead && zw.gen) ? 1 : 0;
			cx.key = zw.dead ? 0xffffffffu : (((unsigned)p.jk << 16) | ((unsigned)(zw.b0 + 1) << 8) | (unsigned)(zw.b1 + 1));
			// bounding box of the shapes that are alive for this slab
			const double bu0 = warp_min_f64(zw.dead ? INFINITY : p.u), bu1 = warp_max_f64(zw.dead ? -INFINITY : p.u);
			const double bv0 = warp_min_f64(zw.dead ? INFINITY : p.v), bv1 = warp_max_f64(zw.dead ? -INFINITY : p.v);

			for (int q = 0; q < n_win; q++) {
				// ---- accumulation window q: r bins [ra, rb] counted from the top ------------------------------------------
				const int rb = P.n_r - 1 - q * W_R, ra = (rb - W_R + 1 > 0) ? rb - W_R + 1 : 0;
				RWindow rw;
				rw.ra = ra;
				rw.lo = P.r2_thr[ra];
				rw.hi = P.r2_thr[rb + 1];
#pragma unroll
				for (int t = 0; t < W_R - 1; t++) rw.thr[t] = (ra + 1 + t <= rb) ? P.r2_thr[ra + 1 + t] : INFINITY;
				const double win_hi = rw.hi;
				cx.rw = rw;
				cx.rb = rb;
				cx.hi_lane = zw.dead ? -1.0 : rw.hi;  // dead lanes never pass the range test
				const double reach_q = sqrt(win_hi) * (1.0 + 1e-9);

				// v regions the warp can reach in this window
				int vr_first = 0, vr_count = n_vr;
				if (n_vr > 1) {
					const double w0 = bv0 - reach_q - 2.0 * eps_v, w1 = bv1 + reach_q + 2.0 * eps_v;
					if (w1 - w0 < L) {
						const int c0 = (int)floor(w0 * P.inv_cv), c1 = (int)floor(w1 * P.inv_cv);
						const int r0 = (int)floor((double)c0 / (double)vr_cells), r1 = (int)floor((double)c1 / (double)vr_cells);
						if (r1 - r0 + 1 < n_vr) {
							vr_first = periodic ? ((r0 % n_vr) + n_vr) % n_vr : (r0 < 0 ? 0 : r0);
							vr_count = periodic ? r1 - r0 + 1 : ((r1 >= n_vr ? n_vr - 1 : r1) - vr_first + 1);
						}
					}
				}

				for (int vri = 0; vri < vr_count; vri++) {
					int g_r = vr_first + vri;
					if (g_r >= n_vr) g_r -= n_vr;
					const int V0 = g_r * vr_cells, V1 = (n_vr > 1) ? V0 + vr_cells - 1 : ncv - 1;
					for (int rbase = 0; rbase < nrows; rbase += 32) {
						// ---- ROUND: up to 32 u rows, one per lane ------------------------------------------------------------
						const int o = rbase + lane;
						int cu = -1;
						if (o < nrows) {
							cu = all_u ? o : ratio * su0 - kk + o;
							if (cu < 0) cu = periodic ? cu + ncu : -1;
							else if (cu >= ncu) cu = periodic ? cu - ncu : -1;
						}
						const long long row = (long long)cu * nz + s;
						// v cells within sqrt(r_hi^2 - d_u^2) of some alive shape of the warp (single precision, rounded outwards), from
						// the GEOMETRIC bounds of the row (cell edges): the range depends on the window and on which lanes are alive,
						// not on the slab or the v region, so it is computed once per (task, window) and reused
						float vmin_f = INFINITY, vmax_f = -INFINITY;
						int cu_code = 0;
						const bool cacheable = (q < 2) && (rbase == 0);
						const unsigned c_al = q == 0 ? c_alive0 : c_alive1;
						if (cacheable && c_al == alive) {
							vmin_f = q == 0 ? c_vmin0 : c_vmin1;
							vmax_f = q == 0 ? c_vmax0 : c_vmax1;
							cu_code = q == 0 ? c_code0 : c_code1;
						} else {
							const double rlo = cu >= 0 ? (double)cu / P.inv_cu - eps_v : INFINITY;
							const double rhi = cu >= 0 ? (double)(cu + 1) / P.inv_cu + eps_v : -INFINITY;
							cu_code = axis_code(bu0, bu1, rlo, rhi);
							for (unsigned mm = alive; mm; mm &= mm - 1u) {
								const int i = __ffs(mm) - 1;
								const double xu = __shfl_sync(0xffffffffu, p.u, i), xv = __shfl_sync(0xffffffffu, p.v, i);
								const double gu = gap(xu, rlo, rhi, cu_code);
								const double g2 = gu * gu * (1.0 - 1e-9);
								if (g2 < win_hi) {
									const double dv = (double)(__fsqrt_ru(__double2float_ru(win_hi - g2)) * 1.000001f) + eps_v;
									vmin_f = fminf(vmin_f, __double2float_rd(xv - dv));
									vmax_f = fmaxf(vmax_f, __double2float_ru(xv + dv));
								}
							}
							if (cacheable) {
								if (q == 0) {
									c_vmin0 = vmin_f;
									c_vmax0 = vmax_f;
									c_code0 = cu_code;
									c_alive0 = alive;
								} else {
									c_vmin1 = vmin_f;
									c_vmax1 = vmax_f;
									c_code1 = cu_code;
									c_alive1 = alive;
								}
							}
						}
						int r_sA = 0, r_eA = 0, r_sB = 0, r_eB = 0, r_lab = -2, r_codes = 0, r_clA = 0, r_clB = 0;
						if (cu >= 0 && vmin_f <= vmax_f) {
							const double vmin = (double)vmin_f, vmax = (double)vmax_f;
							int sa0, sa1, sb0 = 0, sb1 = -1;
							bool none = false;
							if (!periodic) {
								sa0 = cell_index(vmin, P.inv_cv, ncv);
								sa1 = cell_index(vmax, P.inv_cv, ncv);
								none = (vmax < 0.0 || vmin >= L);
							} else if (!(vmax - vmin < L)) {
								sa0 = 0;
								sa1 = ncv - 1;
							} else {
								bool wrapped = false;
								double x0 = vmin, x1 = vmax;
								if (x0 < 0.0) {
									x0 += L;
									wrapped = true;
								}
								if (x1 >= L) {
									x1 -= L;
									wrapped = true;
								}
								const int ca = cell_index(x0, P.inv_cv, ncv), cb_ = cell_index(x1, P.inv_cv, ncv);
								if (!wrapped) {
									sa0 = ca;
									sa1 = cb_;
								} else if (cb_ >= ca - 1) {
									sa0 = 0;
									sa1 = ncv - 1;
								} else {
									sa0 = ca;
									sa1 = ncv - 1;
									sb0 = 0;
									sb1 = cb_;
								}
							}
							sa0 = sa0 > V0 ? sa0 : V0;
							sa1 = sa1 < V1 ? sa1 : V1;
							sb0 = sb0 > V0 ? sb0 : V0;
							sb1 = sb1 < V1 ? sb1 : V1;
							const long long cb0 = row * ncv;
							int cvA = 0, cvB = 0;
							if (!none && sa0 <= sa1) {
								r_sA = (int)a.cell_start[cb0 + sa0];
								r_eA = (int)a.cell_start[cb0 + sa1 + 1];
								r_clA = sa0 | (sa1 << 16);
								cvA = axis_code(bv0, bv1, a.vlo[sa0], a.vhi[sa1]);
							}
							if (!none && sb0 <= sb1) {
								r_sB = (int)a.cell_start[cb0 + sb0];
								r_eB = (int)a.cell_start[cb0 + sb1 + 1];
								r_clB = sb0 | (sb1 << 16);
								cvB = axis_code(bv0, bv1, a.vlo[sb0], a.vhi[sb1]);
							}
							r_codes = cu_code | (cvA << 2) | (cvB << 4);
							if (r_eA > r_sA || r_eB > r_sB) r_lab = a.colreg[row * n_vr + g_r];
						}
						const unsigned m_simple = __ballot_sync(0xffffffffu, r_lab >= 0);
						if (m_simple) process_round2<UNITW, SIG>(&cx, r_sA, r_eA, r_sB, r_eB, r_lab, r_codes, m_simple);
						// ---- (row, region) pairs holding several labels: cell by cell (unaligned grids only) -------------------
						unsigned m_cplx = __ballot_sync(0xffffffffu, r_lab == -1);
						while (m_cplx) {
							const int e = __ffs(m_cplx) - 1;
							m_cplx &= m_cplx - 1u;
							const long long cb0 = ((long long)__shfl_sync(0xffffffffu, cu, e) * nz + s) * ncv;
							const int codes_e = __shfl_sync(0xffffffffu, r_codes, e);
							const int clA = __shfl_sync(0xffffffffu, r_clA, e), clB = __shfl_sync(0xffffffffu, r_clB, e);
							const int okA = __shfl_sync(0xffffffffu, (int)(r_eA > r_sA), e), okB = __shfl_sync(0xffffffffu, (int)(r_eB > r_sB), e);
							for (int piece = 0; piece < 2; piece++) {
								if (!(piece ? okB : okA)) continue;
								const int c0_ = (piece ? clB : clA) & 0xffff, c1_ = (piece ? clB : clA) >> 16;
								const int pc = (codes_e & 3) | (((codes_e >> (2 + 2 * piece)) & 3) << 2);
								for (int cbase = c0_; cbase <= c1_; cbase += 32) {
									int c_s = 0, c_e = 0, c_lab = -2, c_nlab = 0;
									if (cbase + lane <= c1_) {
										c_s = (int)a.cell_start[cb0 + cbase + lane];
										c_e = (int)a.cell_start[cb0 + cbase + lane + 1];
										const CellInfo *cinf = a.cinfo + cb0 + cbase + lane;
										c_nlab = (c_e > c_s) ? cinf->nlab : 0;
										c_lab = (c_nlab == 1) ? cinf->label : -2;
									}
									const unsigned m1 = __ballot_sync(0xffffffffu, c_nlab == 1);
									if (m1) process_round2_cold<UNITW, SIG>(&cx, c_s, c_e, 0, 0, c_lab, pc, m1);
									unsigned mm = __ballot_sync(0xffffffffu, c_nlab > 1);
									while (mm) {
										const int f = __ffs(mm) - 1;
										mm &= mm - 1u;
										int pos = __shfl_sync(0xffffffffu, c_s, f);
										const int end = __shfl_sync(0xffffffffu, c_e, f);
										while (pos < end) {
											const int lb = a.cand_jk[pos];
											int qq = pos + 1;
											while (qq < end && a.cand_jk[qq] == lb) qq++;
											process_round2_cold<UNITW, SIG>(&cx, pos, qq, 0, 0, lb, pc, 1u);
											pos = qq;
										}
									}
								}
							}
						}
					}
				}
				// ---- end of the window: flush what is left in the private slots ------------------------------------------------
				if (cx.cur_label >= 0) {
					cx.binned += flush_slots<UNITW, SIG>(cx.fc, cx.acc, cx.key, zw.dead, cx.pe, p.w, ra, rb, cx.cur_label);
					cx.cur_label = -1;
				}
			}
		}
	}

	}  // next slot

	unsigned long long tested = cx.tested, binned = cx.binned, nan_pairs = cx.nan_pairs;
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

inline int rppi2_prepare(const TiledConfig &cfg, const GridDims &g, const TiledWorkspace &w, cudaStream_t st) {
	MIA_CUDA_CHECK(cudaMemsetAsync(w.vlo, 0xFF, sizeof(double) * g.ncv, st));
	MIA_CUDA_CHECK(cudaMemsetAsync(w.vhi, 0x00, sizeof(double) * g.ncv, st));
	const int64_t nrow = (int64_t)g.ncu * cfg.nz;
	k_row_info<<<(unsigned)((nrow + 127) / 128), 128, 0, st>>>(w.cinfo, nrow, g.ncv, cfg.n_lr, w.colinfo, w.colreg,
															   (unsigned long long *)w.vlo, (unsigned long long *)w.vhi);
	MIA_CUDA_CHECK(cudaGetLastError());
	k_v_envelope<<<1, 32, 0, st>>>(w.vlo, w.vhi, g.ncv);
	MIA_CUDA_CHECK(cudaGetLastError());
	return 0;
}

inline int rppi2_fill_tasks(const TiledArgs &a, const int64_t *prim_cell_start, const int64_t *cell_start, const int32_t *task_off,
							int ncol_s, int nzs, int k, int split, int sym, int32_t *task_col, int64_t *task_first, int32_t *task_n,
							int32_t *task_slab, unsigned long long *task_cost, int32_t *n_tasks, cudaStream_t st) {
	const DevParams &P = a.P;
	k_fill_tasks_rppi2<<<(unsigned)((ncol_s + 127) / 128), 128, 0, st>>>(prim_cell_start, cell_start, task_off, P.ncu, P.ncv, P.ncl,
																		 a.ratio, nzs, split, k, P.periodic, sym, task_col, task_first, task_n,
																		 task_slab, task_cost, n_tasks);
	return (int)cudaGetLastError();
}

template <bool UNITW, bool SIG>
inline int launch_rppi2_t(const TiledArgs &a, int n_ctas, size_t smem, cudaStream_t st) {
	MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rppi2<UNITW, SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_tiled_rppi2<UNITW, SIG><<<n_ctas, TP, smem, st>>>(a);
	return (int)cudaGetLastError();
}

inline int launch_rppi2(const TiledArgs &a, bool unit_w, bool sig, int n_ctas, size_t smem, cudaStream_t st) {
	if (sig) return unit_w ? launch_rppi2_t<true, true>(a, n_ctas, smem, st) : launch_rppi2_t<false, true>(a, n_ctas, smem, st);
	return unit_w ? launch_rppi2_t<true, false>(a, n_ctas, smem, st) : launch_rppi2_t<false, false>(a, n_ctas, smem, st);
}

}  // namespace mia

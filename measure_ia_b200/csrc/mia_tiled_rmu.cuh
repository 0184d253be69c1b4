// mia_tiled_rmu.cuh -- the TILED (r, mu_r) pair kernel for sm_100a (multipoles): fast, bit-reproducible.
//
// Replaces the hot loop of the reference, src/measureia/measure_m_box_jk.py:403-480 (and measure_m_box.py:517-590).
//
// Same machinery as the (r_p, Pi) kernel (mia_tiled.cuh): one thread owns one shape galaxy, one warp owns 32 consecutive
// cell-sorted shape galaxies and is the scheduling unit, position galaxies are streamed into a per-warp double buffer by
// 1-D bulk (TMA) copies behind mbarriers, every thread accumulates into thread-private shared-memory slots with plain
// loads and stores, and all reductions run in a fixed order.  What differs is the geometry:
//   * the search is 3-D.  The grid has cubic cells (columns in the two projected axes x slabs along the line of sight,
//     ALIGNED with the jackknife sub-boxes so that a cell carries one label).  The cells of a column are contiguous in
//     memory, so for every neighbour column the warp streams ONE contiguous range of candidates -- the slabs within
//     sqrt(r_max^2 - d_uv^2) of the warp's shapes -- in chunks of CH_RMU, not cell by cell;
//   * private slots = (W_R r bins) x (all n_mu bins), W_R = 2 if that fits 20 slots, else 1; r bins are visited in
//     windows from the top (the top window holds 80-96 % of the pairs, lower windows re-visit only the nearest cells);
//   * the mu_r bin is floor((mu + 1) n_mu / 2) evaluated with ONE fma on an approximate mu = Pi * rsqrt(r^2) (rel. error
//     3e-16): t = fma(mu, n/2, n/2 + 6145) lands in [4096, 8192), where a double has exactly 40 fractional bits, so the
//     integer part and the distance to the nearest bin edge are read off the bit pattern with integer instructions.
//     Pairs within 1.5e-11 of an edge (and, as in the (r_p, Pi) kernel, pairs with |cos| within 1e-11 of 1) are NOT
//     accumulated by the fast loop: a rare slow path re-evaluates them with the reference's exact operation sequence
//     (mu = Pi / sqrt(r^2), measure_m_box_jk.py:431; thresholds calibrated against numpy).  DD stays bit-exact.
// Candidate chunks carry one jackknife label; the loop order (line-of-sight region, then neighbour columns grouped by
// their projected region) keeps label changes -- each costs a flush of the private slots -- to a handful per task.
#pragma once
#include "mia_tiled.cuh"

namespace mia {

constexpr int NS_RMU = 20;   // private slots per thread
#ifndef MIA_CH_RMU
#define MIA_CH_RMU 64
#endif
constexpr int CH_RMU = MIA_CH_RMU;  // candidates per staged chunk
constexpr unsigned MU_BAND = 16u;   // half-width of the "too close to a mu edge" band in units of 2^-40 bins
constexpr int MAX_NEIGH_RMU = 384;  // neighbour (candidate) columns per task
// Tuning (measured on B200, cfg3: profiles/r01_tuning.md): candidate cells of r_max / DIV, shape columns RATIO x RATIO
// candidate columns wide (so that 32 / HSPLIT consecutive shapes form a compact blob), HSPLIT candidates per warp step.
#ifndef MIA_RMU_DIV
#define MIA_RMU_DIV 6
#endif
#ifndef MIA_RMU_RATIO
#define MIA_RMU_RATIO 2
#endif
#ifndef MIA_RMU_HSPLIT
#define MIA_RMU_HSPLIT 1
#endif
#ifndef MIA_RMU_LMUL
#define MIA_RMU_LMUL 3
#endif
#ifndef MIA_UNROLL_RMU
#define MIA_UNROLL_RMU 2
#endif

inline size_t tiled_rmu_smem_bytes(bool unit_w, bool sig) {
	const size_t fixed = sizeof(Cand) * TW * STAGES * CH_RMU + sizeof(int) * TW * MAX_NEIGH_RMU + 256 + 768;
	const size_t per_slot = (size_t)TP * (8 + 8 + 4 + (unit_w ? 0 : 8) + (sig ? 8 : 0));
	return fixed + per_slot * NS_RMU;
}

// Symmetric (auto-correlation) variant: 48 / 64-byte candidate records (CandSU / CandSW, mia_tiled_rppi2s.cuh), a second
// double2 per slot for the reverse pairs, and only the `ns` slots the configuration uses.
#ifndef MIA_CH_RMU_S
#define MIA_CH_RMU_S 64
#endif
constexpr int CH_RMU_S = MIA_CH_RMU_S;  // largest chunk; smaller ones are used when many slots are needed
constexpr size_t RMU_SYM_SMEM_MAX = 115000;  // two CTAs per SM
inline size_t tiled_rmu_sym_smem_bytes(bool unit_w, int ns, int ch) {
	const size_t rec = unit_w ? 48 : 64;
	const size_t fixed = rec * TW * STAGES * (size_t)ch + sizeof(int) * TW * MAX_NEIGH_RMU + 256 + 768;
	const size_t per_slot = (size_t)TP * (16 + 16 + 4 + (unit_w ? 0 : 8));
	return fixed + per_slot * (size_t)ns;
}
// candidates per staged chunk of the symmetric variant such that two CTAs fit an SM (0: does not fit)
inline int rmu_sym_chunk(bool unit_w, int ns) {
	for (int ch = CH_RMU_S; ch >= 16; ch -= 8)
		if (tiled_rmu_sym_smem_bytes(unit_w, ns, ch) <= RMU_SYM_SMEM_MAX) return ch;
	return 0;
}

// Can the tiled (r, mu_r) kernel take this configuration?  (Declined configurations go to the general kernel.)
inline bool rmu_supported(const mia_params *p, int &w_r) {
	const int n = p->n_2;
	if (n < 1 || n > NS_RMU) return false;
	if (!(p->rp2_cut >= 0.0)) return false;  // r_p = 0 pairs must be excluded by the mask (their e+ is NaN -> 0 in the reference)
	if (!(p->thr2_host[0] == -INFINITY) || !(p->thr2_host[n] == INFINITY)) return false;
	for (int b = 1; b < n; b++) {
		const double nominal = -1.0 + 2.0 * (double)b / (double)n;
		if (!(fabs(p->thr2_host[b] - nominal) <= 1e-13)) return false;
	}
	const double *thr = p->r2_thr_host;
	if (!(thr[0] > 0.0)) return false;
	for (int b = 0; b < p->n_r; b++)
		if (!(thr[b + 1] > thr[b]) || !std::isfinite(thr[b + 1])) return false;
	w_r = (2 * n <= NS_RMU && p->n_r > 1) ? 2 : 1;
	return true;
}


// Grid of the (r, mu_r) kernel: cubic candidate cells of about r_max / DIV, aligned with the jackknife sub-boxes (a cell
// then carries one label) and with the coarser shape columns.
inline bool plan_rmu_grid(const mia_params *p, int n_side, TiledConfig &cfg, int &nc, int &nz, int &k) {
	const double L = p->boxsize, reach = p->r_search * (1.0 + 1e-6);
	int div = env_int("MIA_RMU_DIV", MIA_RMU_DIV), ratio = env_int("MIA_RMU_RATIO", MIA_RMU_RATIO);
	int hs = env_int("MIA_RMU_HSPLIT", MIA_RMU_HSPLIT);
	if (div < 1) div = 1;
	if (ratio < 1) ratio = 1;
	if (hs != 1 && hs != 2 && hs != 4) hs = 1;
	for (;; div--) {
		nc = (int)floor(L / (reach / (double)div));
		if (nc > 512) nc = 512;
		if (nc < 1) nc = 1;
		int rt = ratio;
		while (rt > 1 && nc < 4 * rt) rt--;
		int m = rt;  // nc a multiple of lcm(n_side, ratio) when the box is large enough
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
		const bool all_mode = (2 * k + rt >= nc);
		const long long n_off = all_mode ? (long long)nc * nc : (long long)(2 * k + rt) * (2 * k + rt);
		const unsigned long long nkeys = (unsigned long long)nc * nc * nc * 4ull * (unsigned long long)(p->num_jk > 0 ? p->num_jk : 1);
		if (n_off <= MAX_NEIGH_RMU && nkeys <= (1ull << 31)) {
			cfg.ratio = rt;
			break;
		}
		if (div == 1) return false;
	}
	// slabs thinner than the columns are wide: the streamed range of a column is trimmed to whole slabs, so thin slabs
	// cost nothing (ranges are contiguous) and tighten the culling along the line of sight
	int lmul = env_int("MIA_RMU_LMUL", MIA_RMU_LMUL);
	if (lmul < 1) lmul = 1;
	while (lmul > 1 && ((long long)nc * lmul > 512 ||
						(unsigned long long)nc * nc * nc * lmul * 4ull * (unsigned long long)(p->num_jk > 0 ? p->num_jk : 1) > (1ull << 31)))
		lmul--;
	nz = nc * lmul;
	cfg.hsplit = hs;
	cfg.n_lr = (n_side > 1 && nz % n_side == 0) ? n_side : 1;
	// symmetric auto-correlation variant: the streamed columns (own u rows and k rows ahead) must be unambiguously ahead under
	// the periodic wrap, one candidate per warp step, and two CTAs must fit an SM (only the slots in use are allocated)
	cfg.sym_ok = (2 * (k + cfg.ratio) < nc && hs == 1 &&
				  rmu_sym_chunk(true, cfg.w_r * p->n_2) > 0 && env_int("MIA_RMU_SYM", 1) != 0) ? 1 : 0;
	// (with weights the records and slots are larger: decided again at call time, mia_api.cu)
	return true;
}

// Neighbour enumeration shared by the task-cost kernel and the pair kernel.  Offsets (iu, iv) run over a
// (2k + ratio)^2 window (or the whole grid when that wraps); returns the candidate column or -1.
__device__ __forceinline__ int rmu_neighbour(int o, int su0, int sv0, int ratio, int ncu, int ncv, int k, int periodic,
											 double cs, double reach, int sym = 0) {
	const bool all_u = 2 * k + ratio >= ncu, all_v = 2 * k + ratio >= ncv;
	const int wu = all_u ? ncu : 2 * k + ratio, wv = all_v ? ncv : 2 * k + ratio;
	if (o >= wu * wv) return -1;
	const int iu = o / wv, iv = o - iu * wv;
	const int ou = iu - k, ov = iv - k;  // offsets from the first candidate column of the shape column
	if (sym && ou < 0) return -1;        // symmetric kernel: own u rows and the rows AHEAD only (half-space rule on d_u)
	int nu = all_u ? iu : ratio * su0 + ou, nv = all_v ? iv : ratio * sv0 + ov;
	if (nu < 0) {
		if (!periodic) return -1;
		nu += ncu;
	} else if (nu >= ncu) {
		if (!periodic) return -1;
		nu -= ncu;
	}
	if (nv < 0) {
		if (!periodic) return -1;
		nv += ncv;
	} else if (nv >= ncv) {
		if (!periodic) return -1;
		nv -= ncv;
	}
	if (!all_u && !all_v) {
		const double gu = ou < 0 ? (double)(-ou - 1) : (ou >= ratio ? (double)(ou - ratio) : 0.0);
		const double gv = ov < 0 ? (double)(-ov - 1) : (ov >= ratio ? (double)(ov - ratio) : 0.0);
		if (!((gu * gu + gv * gv) * cs * cs * (1.0 - 1e-6) < reach * reach)) return -1;
	}
	return nu * ncv + nv;
}

__device__ __forceinline__ int rmu_n_offsets(int ratio, int ncu, int ncv, int k) {
	const int wu = (2 * k + ratio >= ncu) ? ncu : 2 * k + ratio, wv = (2 * k + ratio >= ncv) ? ncv : 2 * k + ratio;
	return wu * wv;
}

// One warp task = up to spt consecutive shape galaxies of one shape column; cost = shapes x candidates in reach.
__global__ void k_fill_tasks_rmu(const int64_t *__restrict__ prim_cell_start, const int64_t *__restrict__ cell_start,
								 const int32_t *__restrict__ task_off, int ncu, int ncv, int nz, int ratio, int nzs, int spt,
								 int split, int k, int periodic, int sym, double cs, double reach, int32_t *__restrict__ task_col,
								 int64_t *__restrict__ task_first, int32_t *__restrict__ task_n,
								 int32_t *__restrict__ task_slab, unsigned long long *__restrict__ task_cost,
								 int32_t *__restrict__ n_tasks) {
	const int ncu_s = ncu / ratio, ncv_s = ncv / ratio;
	const int64_t ncol = (int64_t)ncu_s * ncv_s;
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c == 0) n_tasks[0] = task_off[ncol];
	if (c >= ncol) return;
	const int64_t p0 = prim_cell_start[c * nzs], p1 = prim_cell_start[(c + 1) * nzs];
	if (p1 <= p0) return;
	const int su0 = (int)(c / ncv_s), sv0 = (int)(c % ncv_s);
	const int n_off = rmu_n_offsets(ratio, ncu, ncv, k);
	unsigned long long W = 0;
	for (int o = 0; o < n_off; o++) {
		const int nc_ = rmu_neighbour(o, su0, sv0, ratio, ncu, ncv, k, periodic, cs, reach, sym);
		if (nc_ >= 0) W += (unsigned long long)(cell_start[(int64_t)(nc_ + 1) * nz] - cell_start[(int64_t)nc_ * nz]);
	}
	int t = task_off[c];
	for (int64_t p = p0; p < p1; p += spt) {
		const int n = (int)((p1 - p < spt) ? (p1 - p) : spt);
		for (int part = 0; part < split; part++, t++) {  // the same shapes against consecutive parts of the neighbour list
			task_col[t] = (int32_t)c;
			task_first[t] = p;
			task_n[t] = n;
			task_slab[2 * t] = part;
			task_slab[2 * t + 1] = split;
			task_cost[t] = (unsigned long long)(spt / 2 + n / 2) * W / (unsigned long long)split + 1ull;  // (see k_fill_tasks_rppi2)
		}
	}
}

inline int rmu_fill_tasks(const TiledArgs &a, const int64_t *prim_cell_start, const int64_t *cell_start, const int32_t *task_off,
						  int ncol_s, int nzs, int k, int split, int sym, int32_t *task_col, int64_t *task_first, int32_t *task_n,
						  int32_t *task_slab, unsigned long long *task_cost, int32_t *n_tasks, cudaStream_t st) {
	const DevParams &P = a.P;
	const double cs = P.L / P.ncu, reach = sqrt(P.r2_thr[P.n_r]) * (1.0 + 1e-6);
	k_fill_tasks_rmu<<<(unsigned)((ncol_s + 127) / 128), 128, 0, st>>>(prim_cell_start, cell_start, task_off, P.ncu, P.ncv, P.ncl,
																	   a.ratio, nzs, 32 / a.hsplit, split, k, P.periodic, sym, cs, reach, task_col,
																	   task_first, task_n, task_slab, task_cost, n_tasks);
	return (int)cudaGetLastError();
}

// Neighbour candidate columns of a shape column, ordered by jackknife (u, v) region so that candidate labels change
// rarely.  Writes the list to nlist (per-warp shared scratch) and returns its length.
__device__ __noinline__ int build_neighbour_list_rmu(int *nlist, int2 *scratch, int scol, int ratio, int ncu, int ncv, int k,
													 int periodic, int n_side, double cs, double reach, int sym) {
	const int lane = threadIdx.x & 31;
	const int ncv_s = ncv / ratio;
	const int su0 = scol / ncv_s, sv0 = scol % ncv_s;
	const int n_off = rmu_n_offsets(ratio, ncu, ncv, k), n_keys = n_side * n_side;
	__syncwarp();
	for (int o = lane; o < n_off; o += 32) {  // scratch = this warp's (idle) staging buffers
		const int c_ = rmu_neighbour(o, su0, sv0, ratio, ncu, ncv, k, periodic, cs, reach, sym);
		int key = -1;
		if (c_ >= 0) {
			const int nu = c_ / ncv, nv = c_ - nu * ncv;
			key = ((nu * n_side) / ncu) * n_side + (nv * n_side) / ncv;
		}
		scratch[o] = make_int2(c_, key);
	}
	__syncwarp();
	int nn = 0;
	for (int key = 0; key < n_keys; key++) {
		for (int base = 0; base < n_off; base += 32) {
			const int2 ck = (base + lane < n_off) ? scratch[base + lane] : make_int2(-1, -1);
			const bool mine = ck.y == key;
			const unsigned m = __ballot_sync(0xffffffffu, mine);
			if (mine) nlist[nn + __popc(m & ((1u << lane) - 1u))] = ck.x;
			nn += __popc(m);
		}
	}
	__syncwarp();
	return nn;
}

// order-preserving float <-> unsigned maps (warp min / max with one REDUX instruction)
__device__ __forceinline__ unsigned f2ord(float f) {
	const unsigned u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
	return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// The approximate per-pair quantities of the fast loop, in ONE place: the slow path must reproduce the fast loop's
// "suspect" decisions bit for bit.
struct RmuApprox {
	double gp, gc;  // cos 2phi, sin 2phi
	double inv2;    // 2 / r_p^2
	int idx;        // mu bin, clamped to [0, n_mu - 1]
	bool susp;      // too close to a mu edge, or |cos| ~ 1
};

__device__ __forceinline__ RmuApprox rmu_approx(double du, double dv, double dz, double rp2, double s, double a0, double a1,
												 double hn, double tbias, int n_mu) {
	RmuApprox r;
	// 1 / sqrt(s): hardware seed (~2^-22) + one cubically convergent step y (1 + e/2 + 3 e^2 / 8), e = 1 - s y^2
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
	{
		const double e = fma(-(s * y), y, 1.0);
		y = fma(y, fma(0.375, e, 0.5) * e, y);
	}
	const double mu = dz * y;
	const double t = fma(mu, hn, tbias);  // (mu + 1) n_mu / 2 + 6145, rounded to 40 fractional bits
	const unsigned thi = (unsigned)__double2hiint(t), tlo = (unsigned)__double2loint(t);
	const unsigned raw = (thi & 0xFFFFFu) >> 8;  // 2049 + bin
	r.idx = (int)min(max(raw, 2049u), 2048u + (unsigned)n_mu) - 2049;
	const unsigned lo2 = tlo + MU_BAND;
	const unsigned h8 = (thi + (lo2 < MU_BAND ? 1u : 0u)) & 0xFFu;
	const bool susp_mu = (h8 == 0u) && (lo2 < 2u * MU_BAND);
	// e+ / ex as in the (r_p, Pi) kernel
	const double cr = fma(du, a0, __dmul_rn(dv, a1));   // r_p cos(phi)
	const double sr = fma(du, a1, -__dmul_rn(dv, a0));  // r_p sin(phi) (sign irrelevant)
	double z;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(rp2));
	{
		const double e = fma(-rp2, z, 1.0);
		z = fma(z, fma(e, e, e), z);
	}
	const double inv2 = __hiloint2double(__double2hiint(z) + 0x00100000, __double2loint(z));  // 2 / r_p^2
	r.gp = fma(cr * cr, inv2, -1.0);
	r.gc = (cr * fabs(sr)) * inv2;
	r.inv2 = inv2;
	r.susp = susp_mu || (r.gp >= 1.0 - 1e-11);
	return r;
}

// Window of r bins [ra, ra + W_R): limits and the interior threshold on s = r^2, and the r_p^2 cut
struct RmuWindow {
	double lo, hi, thr, cut;
	int ra;
};

// One staged chunk against this thread's shape galaxy.  VAR 0: no periodic image; 1: warp-constant image shifts
// (su, sv, sl); 2: every separation wrapped per pair (the chunk straddles +-L/2 for the warp: tiny boxes only).
// A warp works on hsplit candidates at a time (one per group of 32 / hsplit lanes): cb = address of this lane's first
// candidate, n = number of candidates of this lane, hstep = bytes between them.
template <bool UNITW, bool LOS2, int VAR, bool SIG>
__device__ __forceinline__ bool pair_loop_rmu(uint32_t cb, int n, uint32_t hstep, double L, double halfL, double pu, double pv,
											  double pl, double a0, double a1, double su, double sv, double sl, double w_lo,
											  double w_hi, double w_thr, double w_cut, double hn, double tbias, int n_mu,
											  const PrivAcc &acc) {
	auto wrap = [&](double d) {
		const double c = __hiloint2double(__double2hiint(L) | (__double2hiint(d) & 0x80000000), __double2loint(L));
		return (fabs(d) > halfL) ? __dsub_rn(d, c) : d;  // c = copysign(L, d): measure_m_box_jk.py:419-421
	};
	bool lane_susp = false;
	// current candidate, the next one and the one after it (prefetch distance 2).  The prefetch may run up to two
	// candidates past the end of the chunk: that is still inside this CTA's shared memory and the values are never used.
	double cu, cv, cl, cw, mu_, mv_, ml_, mw_;
	lds_v2(cu, cv, cb);
	lds_v2(cl, cw, cb + 16);
	lds_v2(mu_, mv_, cb + hstep);
	lds_v2(ml_, mw_, cb + hstep + 16);
	uint32_t na = cb + 2u * hstep;
	MIA_UNROLL_PRAGMA(MIA_UNROLL_RMU)
	for (int j = 0; j < n; j++) {
		double nu, nv, nl, nw;
		lds_v2(nu, nv, na);
		lds_v2(nl, nw, na + 16);
		na += hstep;
		double du = __dsub_rn(pu, cu), dv = __dsub_rn(pv, cv), dz = __dsub_rn(pl, cl);  // shape minus position, :418
		if (VAR == 2) {
			du = wrap(du);
			dv = wrap(dv);
			dz = wrap(dz);
		} else if (VAR == 1) {
			du = __dadd_rn(du, su);
			dv = __dadd_rn(dv, sv);
			dz = __dadd_rn(dz, sl);
		}
		const double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dz, dz);
		const double rp2 = __dadd_rn(uu, vv);  // :424 (before the sqrt)
		// r^2 summed over the ORIGINAL columns 0, 1, 2 (:428): (u, v, l) if the line of sight is column 2, else (u, l, v)
		// or (l, u, v), which round identically
		const double s = LOS2 ? __dadd_rn(rp2, ll) : __dadd_rn(__dadd_rn(uu, ll), vv);
		bool ok = (s >= w_lo) && (s < w_hi) && (rp2 > w_cut);
		const RmuApprox ap = rmu_approx(du, dv, dz, rp2, s, a0, a1, hn, tbias, n_mu);
		const int slot = ap.idx + ((s >= w_thr) ? n_mu : 0);
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double s0, s1, sw = 0.0, sq = 0.0;
		lds_v2(s0, s1, acc.a2 + so * 16u);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) sw = lds_f64(acc.aw + so * 8u);
		if (SIG) sq = lds_f64(acc.av + so * 8u);
		lane_susp = lane_susp || (ok && ap.susp);
		ok = ok && !ap.susp;
		double gp = ap.gp, gc = ap.gc;
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64_if(ok, acc.aw + so * 8u, sw + cw);
		}
		if (SIG) sts_f64_if(ok, acc.av + so * 8u, fma(gp, gp, sq));  // (w_D e+)^2: measure_m_box_jk.py:207
		sts_v2_if(ok, acc.a2 + so * 16u, s0 + gp, s1 + gc);
		sts_u32_if(ok, acc.ac + so * 4u, c0 + 1u);
		cu = mu_;
		cv = mv_;
		cl = ml_;
		cw = mw_;
		mu_ = nu;
		mv_ = nv;
		ml_ = nl;
		mw_ = nw;
	}
	return lane_susp;
}

// Rare path: rescan the chunk for the pairs the fast loop skipped and evaluate them exactly as the reference does.
template <bool UNITW, bool LOS2, bool SIG>
__device__ __noinline__ void slow_pairs_rmu(bool lane_susp, uint32_t cb, int n, int periodic, double L, double halfL,
											double pu, double pv, double pl, double a0, double a1, const RmuWindow rw,
											double hi_lane, double hn, double tbias, int n_mu, const double *thr2,
											PrivAcc acc, uint32_t hstep, unsigned long long &nan_pairs) {
	if (!lane_susp) return;
	auto sep = [&](double s_, double c_) {  // measure_m_box_jk.py:418-421
		double d = __dsub_rn(s_, c_);
		if (periodic) {
			if (d > halfL) d = __dsub_rn(d, L);
			if (d < -halfL) d = __dadd_rn(d, L);
		}
		return d;
	};
	for (int j = 0; j < n; j++) {
		double cu, cv, cl, cw;
		lds_v2(cu, cv, cb + (uint32_t)j * hstep);
		lds_v2(cl, cw, cb + (uint32_t)j * hstep + 16);
		const double du = sep(pu, cu), dv = sep(pv, cv), dz = sep(pl, cl);
		const double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dz, dz);
		const double rp2 = __dadd_rn(uu, vv);
		const double s = LOS2 ? __dadd_rn(rp2, ll) : __dadd_rn(__dadd_rn(uu, ll), vv);
		if (!((s >= rw.lo) && (s < hi_lane) && (rp2 > rw.cut))) continue;
		const RmuApprox ap = rmu_approx(du, dv, dz, rp2, s, a0, a1, hn, tbias, n_mu);
		if (!ap.susp) continue;  // the fast loop accumulated this pair
		const double mu = __ddiv_rn(dz, __dsqrt_rn(s));  // :431
		int idx = 0;
		for (int k = 1; k < n_mu; k++) idx += (mu >= thr2[k]) ? 1 : 0;  // :453-460 via the calibrated thresholds
		const double rp = __dsqrt_rn(rp2);
		const double c = __dadd_rn(__dmul_rn(__ddiv_rn(du, rp), a0), __dmul_rn(__ddiv_rn(dv, rp), a1));  // :432-436
		double gp = 0.0, gc = 0.0;
		if (fabs(c) <= 1.0) shape_projection(c, gp, gc);
		else nan_pairs++;  // arccos -> NaN -> e+ = ex = 0, the pair still counts (:437-438)
		const int slot = idx + ((s >= rw.thr) ? n_mu : 0);
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

// Flush: fixed-order warp reduction of the private slots into this warp's accumulator copy in HBM.  All lanes share the
// slot -> bin map (slot = r_offset * n_mu + mu bin); lanes are grouped by the jackknife label of their shape galaxy.
template <bool UNITW, bool SIG>
__device__ __noinline__ unsigned flush_slots_rmu(const FlushCtx &fc, PrivAcc acc, int jkS, bool dead, double pe, double pw,
												 int ra, int rb, int n_mu, int ns, int jkD) {
	const int lane = threadIdx.x & 31;
	unsigned binned = 0;
	unsigned todo = __ballot_sync(0xffffffffu, !dead);
	while (todo) {
		const int leader = __ffs(todo) - 1;
		const int k = __shfl_sync(0xffffffffu, jkS, leader);
		const unsigned grp = __ballot_sync(0xffffffffu, jkS == k) & todo;
		const bool in = (grp >> lane) & 1u;
		unsigned tot_cnt = 0;
		double tot_sp = 0.0, tot_sc = 0.0, tot_dw = 0.0, tot_sq = 0.0;
#pragma unroll 1
		for (int sl = 0; sl < ns; sl++) {
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
		if (lane < ns && tot_cnt) {
			const int roff = lane / n_mu, b2 = lane - roff * n_mu;
			const int rbin = ra + roff;
			if (rbin > rb) {
				atomicExch(&fc.flags[1], 1);
			} else {
				// all loads before the first store: one memory latency per flush group instead of seven
				const size_t bin = (size_t)rbin * fc.n_2 + b2;
				const size_t ia = (size_t)k * fc.nb + bin;
				const bool has_b = fc.num_jk > 0 && jkD != k;
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
#pragma unroll 1
	for (int sl = 0; sl < ns; sl++) {
		sts_v2(acc.a2 + (uint32_t)sl * TP * 16u, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)sl * TP * 8u, 0.0);
		if (SIG) sts_f64(acc.av + (uint32_t)sl * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)sl * TP * 4u, 0u);
	}
	return binned;
}

// ------------------------------------------------------------------------------------------------------------------
// SYMMETRIC variant (auto-correlations; see mia_tiled_rppi2s.cuh for the idea): every unordered pair is visited once, by the
// galaxy that sees the other one ahead along u (wrapped d_u < 0); the separation, r^2, the range tests, the r bin, the
// reciprocals and mu are shared, the reverse pair has mu -> -mu (mirrored mu bin: the SAME private slot, second double2) and
// its own projected shape (the candidate record carries its axis and w * e).  Pairs near a mu edge or with |cos| ~ 1 in
// either ordering, and exact ties d_u == 0, go to an exact path that adds straight into the slot's accumulator copy.
// ------------------------------------------------------------------------------------------------------------------
template <bool UNITW>
struct RmuRec {
	typedef CandSW type;
};
template <>
struct RmuRec<true> {
	typedef CandSU type;
};

struct RmuPairS {
	RmuApprox ap;      // forward e+ / ex, mu bin, suspicion flag
	double gpr, gcr;   // reverse e+ / ex (unweighted)
	unsigned bad;      // some ordering needs the exact path
};
__device__ __forceinline__ RmuPairS rmu_pair_sym(double du, double dv, double dz, double rp2, double s, double a0, double a1,
												 double b0, double b1, double hn, double tbias, int n_mu) {
	RmuPairS r;
	r.ap = rmu_approx(du, dv, dz, rp2, s, a0, a1, hn, tbias, n_mu);
	const double crr = fma(du, b0, __dmul_rn(dv, b1));   // reverse: separation -sep, cos = -crr / r_p
	const double srr = fma(du, b1, -__dmul_rn(dv, b0));
	r.gpr = fma(crr * crr, r.ap.inv2, -1.0);
	r.gcr = -((crr * fabs(srr)) * r.ap.inv2);
	r.bad = (unsigned)r.ap.susp | (unsigned)(r.gpr >= 1.0 - 1e-11);
	return r;
}

template <bool UNITW, bool LOS2, int VAR>
__device__ __forceinline__ bool pair_loop_rmu_sym(uint32_t cb, int n, double L, double halfL, double pu, double pv, double pl,
												  double a0, double a1, double su, double sv, double sl, double w_lo, double w_hi,
												  double w_thr, double w_cut, double hn, double tbias, int n_mu, const PrivAcc &acc,
												  uint32_t ar_off) {
	constexpr uint32_t REC = (uint32_t)sizeof(typename RmuRec<UNITW>::type);
	auto wrap = [&](double d) {
		const double c = __hiloint2double(__double2hiint(L) | (__double2hiint(d) & 0x80000000), __double2loint(L));
		return (fabs(d) > halfL) ? __dsub_rn(d, c) : d;
	};
	unsigned lane_susp = 0u;
	double cu, cv, cl, we, b0, b1, cw = 1.0, pad_;
	lds_v2(cu, cv, cb);
	lds_v2(cl, we, cb + 16);
	lds_v2(b0, b1, cb + 32);
	if (!UNITW) lds_v2(cw, pad_, cb + 48);
	uint32_t na = cb + REC;
	MIA_UNROLL_PRAGMA(MIA_UNROLL_RMU)
	for (int j = 0; j < n; j++) {
		double nu, nv, nl, nwe, nb0, nb1, nw = 1.0;
		lds_v2(nu, nv, na);
		lds_v2(nl, nwe, na + 16);
		lds_v2(nb0, nb1, na + 32);
		if (!UNITW) lds_v2(nw, pad_, na + 48);
		na += REC;
		double du = __dsub_rn(pu, cu), dv = __dsub_rn(pv, cv), dz = __dsub_rn(pl, cl);  // shape minus position, :418
		if (VAR == 2) {
			du = wrap(du);
			dv = wrap(dv);
			dz = wrap(dz);
		} else if (VAR == 1) {
			du = __dadd_rn(du, su);
			dv = __dadd_rn(dv, sv);
			dz = __dadd_rn(dz, sl);
		}
		const double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dz, dz);
		const double rp2 = __dadd_rn(uu, vv);
		const double s = LOS2 ? __dadd_rn(rp2, ll) : __dadd_rn(__dadd_rn(uu, ll), vv);
		const unsigned inr = (s >= w_lo) & (s < w_hi) & (rp2 > w_cut);
		const RmuPairS pr = rmu_pair_sym(du, dv, dz, rp2, s, a0, a1, b0, b1, hn, tbias, n_mu);
		const int slot = pr.ap.idx + ((s >= w_thr) ? n_mu : 0);
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double f0, f1, r0, r1, sw = 0.0;
		lds_v2(f0, f1, acc.a2 + so * 16u);
		lds_v2(r0, r1, acc.a2 + so * 16u + ar_off);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) sw = lds_f64(acc.aw + so * 8u);
		const int dh = __double2hiint(du);
		const unsigned neg = (unsigned)(dh < 0);
		const unsigned tie = (unsigned)((((unsigned)dh << 1) | (unsigned)__double2loint(du)) == 0u);
		lane_susp |= inr & ((neg & pr.bad) | tie);
		const bool ok = (inr & neg & (pr.bad ^ 1u)) != 0u;
		double gp = pr.ap.gp, gc = pr.ap.gc;
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64_if(ok, acc.aw + so * 8u, sw + cw);
		}
		sts_v2_if(ok, acc.a2 + so * 16u, f0 + gp, f1 + gc);
		sts_v2_if(ok, acc.a2 + so * 16u + ar_off, fma(pr.gpr, we, r0), fma(pr.gcr, we, r1));
		sts_u32_if(ok, acc.ac + so * 4u, c0 + 1u);
		cu = nu;
		cv = nv;
		cl = nl;
		we = nwe;
		b0 = nb0;
		b1 = nb1;
		cw = nw;
	}
	return lane_susp != 0u;
}

// rows A[jk_shape] (+ B[jk_pos] when the position galaxy lies in another region) += one (count, sum w, sum e+, sum ex);
// loads before stores
__device__ __forceinline__ void add_rows_fc(const FlushCtx &fc, size_t bin, int jk_shape, int jk_pos, unsigned long long cnt, double dw,
											double sp, double sc) {
	const size_t ia = (size_t)jk_shape * fc.nb + bin;
	const bool has_b = fc.num_jk > 0 && jk_pos != jk_shape;
	const size_t ib = has_b ? (size_t)(fc.J + jk_pos) * fc.nb + bin : ia;
	const unsigned long long c_a = fc.pcnt[ia], c_b = fc.pcnt[ib];
	const double d_a = fc.pddw[ia], p_a = fc.psp[ia], x_a = fc.psc[ia], d_b = fc.pddw[ib], p_b = fc.psp[ib];
	fc.pcnt[ia] = c_a + cnt;
	fc.pddw[ia] = d_a + dw;
	fc.psp[ia] = p_a + sp;
	fc.psc[ia] = x_a + sc;
	if (has_b) {
		fc.pcnt[ib] = c_b + cnt;
		fc.pddw[ib] = d_b + dw;
		fc.psp[ib] = p_b + sp;
	}
}

// Rare path of the symmetric variant: the pairs the fast loop left out, one lane at a time, both orderings evaluated exactly
// as the reference does (mu = Pi / sqrt(r^2), cos, NaN rule: measure_m_box_jk.py:431-438), added to the slot's copy.
template <bool UNITW, bool LOS2>
__device__ __noinline__ void slow_pairs_rmu_sym(bool lane_susp, uint32_t cb, int n, int periodic, double L, double halfL,
												double pu, double pv, double pl, double a0, double a1, double pe, double pw, int jkS,
												int jkD, const RmuWindow rw, double w_hi, double hn, double tbias, int n_mu,
												const double *thr2, const FlushCtx fc, unsigned long long &nan_pairs,
												unsigned long long &binned) {
	constexpr uint32_t REC = (uint32_t)sizeof(typename RmuRec<UNITW>::type);
	const int lane = threadIdx.x & 31;
	auto sep = [&](double s_, double c_) {  // measure_m_box_jk.py:418-421
		double d = __dsub_rn(s_, c_);
		if (periodic) {
			if (d > halfL) d = __dsub_rn(d, L);
			if (d < -halfL) d = __dadd_rn(d, L);
		}
		return d;
	};
	for (unsigned m = __ballot_sync(0xffffffffu, lane_susp); m; m &= m - 1u) {
		if (lane == __ffs(m) - 1) {
			for (int j = 0; j < n; j++) {
				double cu, cv, cl, we, b0, b1, cw = 1.0, pad_;
				const uint32_t ca = cb + (uint32_t)j * REC;
				lds_v2(cu, cv, ca);
				lds_v2(cl, we, ca + 16);
				lds_v2(b0, b1, ca + 32);
				if (!UNITW) lds_v2(cw, pad_, ca + 48);
				const double du = sep(pu, cu), dv = sep(pv, cv), dz = sep(pl, cl);
				const double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dz, dz);
				const double rp2 = __dadd_rn(uu, vv);
				const double s = LOS2 ? __dadd_rn(rp2, ll) : __dadd_rn(__dadd_rn(uu, ll), vv);
				if (!((s >= rw.lo) && (s < w_hi) && (rp2 > rw.cut))) continue;
				// this galaxy takes the pair iff the other one is ahead: (d_u, d_v, d_z) lexicographically negative
				if (!(du < 0.0 || (du == 0.0 && (dv < 0.0 || (dv == 0.0 && dz < 0.0))))) continue;
				const RmuPairS pr = rmu_pair_sym(du, dv, dz, rp2, s, a0, a1, b0, b1, hn, tbias, n_mu);
				if (!(du == 0.0 || pr.bad)) continue;  // the fast loop accumulated this pair
				const double r = __dsqrt_rn(s), rp = __dsqrt_rn(rp2);
				const int rbin = rw.ra + ((s >= rw.thr) ? 1 : 0);
				const double ww = pw * cw;
				for (int dir = 0; dir < 2; dir++) {
					const double mu = __ddiv_rn(dir ? -dz : dz, r);  // :431; the reverse pair has sep_ji = -sep_ij exactly
					int idx = 0;
					for (int k = 1; k < n_mu; k++) idx += (mu >= thr2[k]) ? 1 : 0;
					const double x0 = dir ? -du : du, x1 = dir ? -dv : dv;
					const double c = dir ? __dadd_rn(__dmul_rn(__ddiv_rn(x0, rp), b0), __dmul_rn(__ddiv_rn(x1, rp), b1))
										 : __dadd_rn(__dmul_rn(__ddiv_rn(x0, rp), a0), __dmul_rn(__ddiv_rn(x1, rp), a1));
					double gp = 0.0, gc = 0.0;
					if (fabs(c) <= 1.0) shape_projection(c, gp, gc);
					else nan_pairs++;
					const double amp = dir ? pw * we : pe * cw;  // w_D w_S e_S
					add_rows_fc(fc, (size_t)rbin * fc.n_2 + idx, dir ? jkD : jkS, dir ? jkS : jkD, 1ull, ww, gp * amp, gc * amp);
					binned++;
				}
			}
		}
		__syncwarp();
	}
}

// Flush of the symmetric variant: as flush_slots_rmu, then -- after a warp barrier, because forward and reverse bins of
// different lanes coincide -- the reverse sums go to the mirrored mu bin of rows A[label of the chunk] / B[label of the lane].
template <bool UNITW>
__device__ __noinline__ unsigned flush_slots_rmu_sym(const FlushCtx &fc, PrivAcc acc, uint32_t ar_off, int jkS, bool dead, double pe,
													 double pw, int ra, int rb, int n_mu, int ns, int jkD) {
	const int lane = threadIdx.x & 31;
	unsigned binned = 0;
	unsigned todo = __ballot_sync(0xffffffffu, !dead);
	while (todo) {
		const int leader = __ffs(todo) - 1;
		const int k = __shfl_sync(0xffffffffu, jkS, leader);
		const unsigned grp = __ballot_sync(0xffffffffu, jkS == k) & todo;
		const bool in = (grp >> lane) & 1u;
		unsigned tot_cnt = 0;
		double tf_p = 0.0, tf_c = 0.0, tr_p = 0.0, tr_c = 0.0, tot_dw = 0.0;
#pragma unroll 1
		for (int sl = 0; sl < ns; sl++) {
			const uint32_t so = (uint32_t)sl * TP;
			const unsigned c = in ? lds_u32(acc.ac + so * 4u) : 0u;
			const unsigned csum = __reduce_add_sync(0xffffffffu, c);
			if (csum == 0u) continue;
			double v0 = 0.0, v1 = 0.0, u0 = 0.0, u1 = 0.0;
			if (in) {
				lds_v2(v0, v1, acc.a2 + so * 16u);
				lds_v2(u0, u1, acc.a2 + so * 16u + ar_off);
			}
			const double xf = warp_sum(v0 * pe), yf = warp_sum(v1 * pe);
			const double xr = warp_sum(u0 * pw), yr = warp_sum(u1 * pw);
			const double zs = UNITW ? (double)csum : warp_sum(in ? lds_f64(acc.aw + so * 8u) * pw : 0.0);
			if (lane == sl) {
				tot_cnt = csum;
				tf_p = xf;
				tf_c = yf;
				tr_p = xr;
				tr_c = yr;
				tot_dw = zs;
			}
		}
		const int roff = lane / n_mu, b2 = lane - roff * n_mu;
		const int rbin = ra + roff;
		const bool mine = lane < ns && tot_cnt;
		if (mine && rbin > rb) atomicExch(&fc.flags[1], 1);
		const bool go = mine && rbin <= rb;
		if (go) add_rows_fc(fc, (size_t)rbin * fc.n_2 + b2, k, jkD, tot_cnt, tot_dw, tf_p, tf_c);
		__syncwarp();
		if (go) {
			add_rows_fc(fc, (size_t)rbin * fc.n_2 + (n_mu - 1 - b2), jkD, k, tot_cnt, tot_dw, tr_p, tr_c);
			binned += 2u * tot_cnt;
		}
		__syncwarp();
		todo &= ~grp;
	}
#pragma unroll 1
	for (int sl = 0; sl < ns; sl++) {
		sts_v2(acc.a2 + (uint32_t)sl * TP * 16u, 0.0, 0.0);
		sts_v2(acc.a2 + (uint32_t)sl * TP * 16u + ar_off, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)sl * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)sl * TP * 4u, 0u);
	}
	return binned;
}

__device__ __forceinline__ double warp_min_f64(double x) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
	return x;
}
__device__ __forceinline__ double warp_max_f64(double x) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
	return x;
}

// Everything the chunk consumer needs for one (task, window).  Lives in the kernel's local memory; process_round() --
// a __noinline__ function with its own register allocation -- loads it once per ROUND of up to 32 column descriptors.
struct RmuCtx {
	// per lane
	double pu, pv, pl, a0, a1, hi_lane, pe, pw;
	int jk, dead;
	// per (task, window)
	double L, halfL, lo, hi, thr, cut, hn, tbias;
	int n_mu, ns, ra, rb, periodic, hlog, half;
	uint32_t a2, aw, ac, av, ar_off, ring_u32, hstep;
	int ch;  // symmetric variant: candidates per staged chunk
	const unsigned char *cand;
	unsigned char *ring;
	uint64_t *full;
	const double *thr2;
	FlushCtx fc;
	// mutable: stream state, current candidate label, statistics
	uint32_t phase0, phase1;
	int st_issue, cur_label;
	unsigned long long tested, binned, nan_pairs;
};

// Image codes of a chunk: per axis 0 = no image, 1 = every pair wraps down (d -= L), 2 = up (d += L), 3 = the chunk
// straddles +-L/2 for the warp (wrap per pair).  codes = cu | cv << 2 | cl(piece A) << 4 | cl(piece B) << 6.
__device__ __forceinline__ double code_shift(int code, double L) { return code == 1 ? -L : (code == 2 ? L : 0.0); }

// Consume one round: lane e (bit e of mask) holds a column descriptor = up to two contiguous candidate ranges [sA, eA),
// [sB, eB) with ONE jackknife label `lab` and warp-constant image codes.  Ranges are cut into chunks of <= CH_RMU,
// streamed through the warp's double buffer (bulk copy of chunk k+1 in flight while chunk k is processed).
template <bool UNITW, bool LOS2, bool SIG, bool SYM>
__device__ __noinline__ void process_round(RmuCtx *cx, int sA, int eA, int sB, int eB, int lab, int codes, unsigned mask) {
	const int lane = threadIdx.x & 31;
	const double L = cx->L, halfL = cx->halfL, pu = cx->pu, pv = cx->pv, pl = cx->pl, a0 = cx->a0, a1 = cx->a1;
	const double hi_lane = cx->hi_lane, hn = cx->hn, tbias = cx->tbias;
	RmuWindow rw;
	rw.lo = cx->lo;
	rw.hi = cx->hi;
	rw.thr = cx->thr;
	rw.cut = cx->cut;
	rw.ra = cx->ra;
	const int n_mu = cx->n_mu, hlog = cx->hlog, half = cx->half;
	const bool dead = cx->dead != 0;
	PrivAcc acc;
	acc.a2 = cx->a2;
	acc.aw = cx->aw;
	acc.ac = cx->ac;
	acc.av = cx->av;
	constexpr uint32_t REC = SYM ? (uint32_t)sizeof(typename RmuRec<UNITW>::type) : (uint32_t)sizeof(Cand);
	const int CHR = SYM ? cx->ch : CH_RMU;  // candidates per staged chunk
	const uint32_t ring_u32 = cx->ring_u32, hstep = cx->hstep, ar_off = cx->ar_off;
	const unsigned char *cand = cx->cand;
	unsigned char *ring = cx->ring;
	uint64_t *full = cx->full;
	uint32_t phase0 = cx->phase0, phase1 = cx->phase1;
	int st_issue = cx->st_issue, cur_label = cx->cur_label;
	unsigned long long tested = 0, nan_pairs = 0, tested_extra = 0;  // tested_extra: ordered pairs binned by the exact path
	unsigned binned = 0;

	int pend_n = 0, pend_st = 0, pend_label = -1, pend_codes = 0;
	auto consume = [&]() {
		if (pend_label != cur_label) {  // candidates of another jackknife region: flush the private slots
			if (cur_label >= 0) {
				if (SYM)
					binned += flush_slots_rmu_sym<UNITW>(cx->fc, acc, ar_off, cx->jk, dead, cx->pe, cx->pw, rw.ra, cx->rb, n_mu, cx->ns, cur_label);
				else
					binned += flush_slots_rmu<UNITW, SIG>(cx->fc, acc, cx->jk, dead, cx->pe, cx->pw, rw.ra, cx->rb, n_mu, cx->ns, cur_label);
			}
			cur_label = pend_label;
		}
		const int cu_ = pend_codes & 3, cv_ = (pend_codes >> 2) & 3, cl_ = (pend_codes >> 4) & 3;
		if (pend_st == 0) {
			mbar_wait(&full[0], phase0);
			phase0 ^= 1u;
		} else {
			mbar_wait(&full[1], phase1);
			phase1 ^= 1u;
		}
		const int n_mine = (pend_n - half + (1 << hlog) - 1) >> hlog;  // candidates half, half + hsplit, ... of the chunk
		if (!dead) tested += (unsigned long long)n_mine;
		const uint32_t cb = ring_u32 + (uint32_t)pend_st * (uint32_t)(CHR * REC) + (uint32_t)half * REC;
		bool susp;
		if (SYM) {  // (hsplit == 1: half = 0, n_mine = pend_n)
			if (cu_ == 3 || cv_ == 3 || cl_ == 3)
				susp = pair_loop_rmu_sym<UNITW, LOS2, 2>(cb, n_mine, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, 0.0, rw.lo, hi_lane, rw.thr, rw.cut,
														 hn, tbias, n_mu, acc, ar_off);
			else if (pend_codes & 63)
				susp = pair_loop_rmu_sym<UNITW, LOS2, 1>(cb, n_mine, L, halfL, pu, pv, pl, a0, a1, code_shift(cu_, L), code_shift(cv_, L),
														 code_shift(cl_, L), rw.lo, hi_lane, rw.thr, rw.cut, hn, tbias, n_mu, acc, ar_off);
			else
				susp = pair_loop_rmu_sym<UNITW, LOS2, 0>(cb, n_mine, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, 0.0, rw.lo, hi_lane, rw.thr, rw.cut,
														 hn, tbias, n_mu, acc, ar_off);
			if (__any_sync(0xffffffffu, susp)) {
				unsigned long long b_ = 0ull;
				slow_pairs_rmu_sym<UNITW, LOS2>(susp, cb, n_mine, cx->periodic, L, halfL, pu, pv, pl, a0, a1, cx->pe, cx->pw, cx->jk,
												cur_label, rw, rw.hi, hn, tbias, n_mu, cx->thr2, cx->fc, nan_pairs, b_);
				tested_extra += b_;
			}
			__syncwarp();
			return;
		}
		if (cu_ == 3 || cv_ == 3 || cl_ == 3)
			susp = pair_loop_rmu<UNITW, LOS2, 2, SIG>(cb, n_mine, hstep, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, 0.0, rw.lo, hi_lane, rw.thr,
												 rw.cut, hn, tbias, n_mu, acc);
		else if (pend_codes & 63)
			susp = pair_loop_rmu<UNITW, LOS2, 1, SIG>(cb, n_mine, hstep, L, halfL, pu, pv, pl, a0, a1, code_shift(cu_, L),
												 code_shift(cv_, L), code_shift(cl_, L), rw.lo, hi_lane, rw.thr, rw.cut, hn, tbias,
												 n_mu, acc);
		else
			susp = pair_loop_rmu<UNITW, LOS2, 0, SIG>(cb, n_mine, hstep, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, 0.0, rw.lo, hi_lane, rw.thr,
												 rw.cut, hn, tbias, n_mu, acc);
		if (__any_sync(0xffffffffu, susp))
			slow_pairs_rmu<UNITW, LOS2, SIG>(susp, cb, n_mine, cx->periodic, L, halfL, pu, pv, pl, a0, a1, rw, hi_lane, hn, tbias, n_mu,
										cx->thr2, acc, hstep, nan_pairs);
		__syncwarp();  // every lane is done with the stage before it is refilled
	};

	while (mask) {
		const int e = __ffs(mask) - 1;
		mask &= mask - 1u;
		const int d_lab = __shfl_sync(0xffffffffu, lab, e), d_codes = __shfl_sync(0xffffffffu, codes, e);
		const int d_sA = __shfl_sync(0xffffffffu, sA, e), d_eA = __shfl_sync(0xffffffffu, eA, e);
		const int d_sB = __shfl_sync(0xffffffffu, sB, e), d_eB = __shfl_sync(0xffffffffu, eB, e);
#pragma unroll 1
		for (int piece = 0; piece < 2; piece++) {
			int s = piece ? d_sB : d_sA;
			const int en = piece ? d_eB : d_eA;
			const int pc = (d_codes & 15) | (((d_codes >> (4 + 2 * piece)) & 3) << 4);
			while (s < en) {
				const int rest = en - s, nch = (rest + CHR - 1) / CHR;  // equal chunks (66 -> 33 + 33, not 64 + 2)
				const int n = (rest + nch - 1) / nch;
				if (lane == 0) {
					const uint32_t bytes = (uint32_t)n * REC;
					mbar_expect_tx(&full[st_issue], bytes);
					bulk_load(ring + (size_t)st_issue * CHR * REC, cand + (size_t)s * REC, bytes, &full[st_issue]);
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
	cx->binned += binned + tested_extra;
	cx->nan_pairs += nan_pairs;
}

template <bool UNITW, bool LOS2, bool SIG, bool SYM>
__global__ void __launch_bounds__(TP, SYM ? 2 : 3) k_tiled_rmu(const TiledArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	const DevParams &P = a.P;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb = P.n_r * P.n_2;
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	const int periodic = P.periodic;
	const double L = P.L, halfL = P.halfL;
	const int n_mu = P.n_2, w_r = a.w_r, ns = w_r * n_mu;
	const int nz = a.nz;
	// a warp works on spt = 32 / hsplit shape galaxies; lane group `half` takes every hsplit-th candidate of a chunk
	const int hsplit = a.hsplit, spt = 32 / hsplit;
	const int sidx = lane & (spt - 1);

	// ---- shared memory carve-up ------------------------------------------------------------------------------------
	constexpr uint32_t REC = SYM ? (uint32_t)sizeof(typename RmuRec<UNITW>::type) : (uint32_t)sizeof(Cand);
	const int CHR = SYM ? a.ch_sym : CH_RMU;
	const int ns_alloc = SYM ? ns : NS_RMU;  // slots carved out of shared memory
	unsigned char *ring = smem;  // [warp][stage][CHR] candidate records
	int *nlist_all = reinterpret_cast<int *>(smem + (size_t)REC * TW * STAGES * CHR);
	uint64_t *full = reinterpret_cast<uint64_t *>(nlist_all + TW * MAX_NEIGH_RMU);  // [warp][stage]
	double *thr2_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(full) + 256);
	unsigned char *accbase = reinterpret_cast<unsigned char *>(thr2_s) + 768;
	const uint32_t acc_u32 = smem_u32(accbase);
	unsigned char *my_ring = ring + (size_t)warp * STAGES * CHR * REC;
	int *nlist = nlist_all + warp * MAX_NEIGH_RMU;

	RmuCtx cx;
	// layout: a2 [slot][thread] 16 B | (SYM) reverse double2 | (weights) aw 8 B | (SIG) av 8 B | ac 4 B
	const uint32_t f_bytes = SYM ? 32u : 16u;
	cx.a2 = acc_u32 + (uint32_t)tid * 16u;
	cx.ar_off = (uint32_t)ns_alloc * TP * 16u;
	cx.aw = acc_u32 + (uint32_t)ns_alloc * TP * f_bytes + (uint32_t)tid * 8u;
	cx.av = acc_u32 + (uint32_t)ns_alloc * TP * (f_bytes + (UNITW ? 0u : 8u)) + (uint32_t)tid * 8u;
	cx.ac = acc_u32 + (uint32_t)ns_alloc * TP * (f_bytes + (UNITW ? 0u : 8u) + (SIG ? 8u : 0u)) + (uint32_t)tid * 4u;
	cx.ring_u32 = smem_u32(my_ring);
	cx.ring = my_ring;
	cx.full = full + warp * STAGES;
	cx.cand = reinterpret_cast<const unsigned char *>(a.cand);
	cx.thr2 = thr2_s;
	cx.hstep = (uint32_t)hsplit * REC;
	cx.ch = CHR;
	cx.hlog = hsplit == 4 ? 2 : (hsplit == 2 ? 1 : 0);
	cx.half = lane / spt;
	cx.L = L;
	cx.halfL = halfL;
	cx.periodic = periodic;
	cx.n_mu = n_mu;
	cx.ns = ns;
	cx.hn = 0.5 * (double)n_mu;
	cx.tbias = cx.hn + 6145.0;
	cx.cut = P.rp2_cut;
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
#pragma unroll 1
	for (int s = 0; s < ns_alloc; s++) {
		sts_v2(cx.a2 + (uint32_t)s * TP * 16u, 0.0, 0.0);
		if (SYM) sts_v2(cx.a2 + (uint32_t)s * TP * 16u + cx.ar_off, 0.0, 0.0);
		if (!UNITW) sts_f64(cx.aw + (uint32_t)s * TP * 8u, 0.0);
		if (SIG) sts_f64(cx.av + (uint32_t)s * TP * 8u, 0.0);
		sts_u32(cx.ac + (uint32_t)s * TP * 4u, 0u);
	}
	__syncthreads();  // the only CTA-wide synchronisation

	// ---- this warp's share of the tasks (same rule as the (r_p, Pi) kernel) -------------------------------------------
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

	const double TN = P.r2_thr[P.n_r];
	const double cs = L / P.ncu, reach = sqrt(TN) * (1.0 + 1e-6);
	const int n_win = (P.n_r + w_r - 1) / w_r;
	// line-of-sight regions: when the slabs are aligned with the jackknife sub-boxes, candidates are visited region by
	// region so that the label of consecutive chunks changes as rarely as possible
	const int n_lr = a.n_lr;
	const int lr_cells = nz / n_lr;
	const double eps_l = 1e-9 * L;

	// image code of an axis for the whole warp: shapes in [b0, b1], candidates in [cmin, cmax]; fl(shape - candidate) is
	// monotone in both, so the two extreme differences bracket every pair's
	auto axis_code = [&](double b0, double b1, double cmin, double cmax) -> int {
		if (!periodic) return 0;
		const double lo = __dsub_rn(b0, cmax), hi = __dsub_rn(b1, cmin);
		if (lo >= -halfL && hi <= halfL) return 0;
		if (lo > halfL) return 1;   // every pair wraps down: sep -= L (measure_m_box_jk.py:420)
		if (hi < -halfL) return 2;  // every pair wraps up (:421)
		return 3;
	};
	// distance of the point x from the interval [cmin, cmax] under the image of `code` (3: nearest of the three images)
	auto gap = [&](double x, double cmin, double cmax, int code) -> double {
		if (code == 3) {
			double g = fmax(0.0, fmax(cmin - x, x - cmax));
			g = fmin(g, fmax(0.0, fmax((cmin + L) - x, x - (cmax + L))));
			return fmin(g, fmax(0.0, fmax((cmin - L) - x, x - (cmax - L))));
		}
		const double sh = code == 1 ? L : (code == 2 ? -L : 0.0);  // candidates seen at c + sh
		return fmax(0.0, fmax((cmin + sh) - x, x - (cmax + sh)));
	};

	for (int task = task0; task < task1; task++) {
		const int col = a.task_col[task];
		const int np = a.task_n[task];
		const bool dead = sidx >= np;
		Prim p;
		if (!dead) {
			p = a.prim[a.task_first[task] + sidx];
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
		cx.jk = p.jk;
		cx.dead = dead ? 1 : 0;
		int nn = build_neighbour_list_rmu(nlist, reinterpret_cast<int2 *>(my_ring), col, a.ratio, P.ncu, P.ncv, P.ku, periodic,
										  a.n_side, cs, reach, SYM ? 1 : 0);
		// when tasks are scarce (small catalogues, many GPUs) a task covers one of `nparts` consecutive parts of the list
		int noff = 0;
		{
			const int part = a.task_slab[2 * task], nparts = a.task_slab[2 * task + 1];
			if (nparts > 1) {
				noff = (int)((long long)nn * part / nparts);
				nn = (int)((long long)nn * (part + 1) / nparts) - noff;
			}
		}
		// bounding box of the warp's shape galaxies
		const double bu0 = warp_min_f64(dead ? INFINITY : p.u), bu1 = warp_max_f64(dead ? -INFINITY : p.u);
		const double bv0 = warp_min_f64(dead ? INFINITY : p.v), bv1 = warp_max_f64(dead ? -INFINITY : p.v);
		const double sl0 = warp_min_f64(dead ? INFINITY : p.l), sl1 = warp_max_f64(dead ? -INFINITY : p.l);

		for (int q = 0; q < n_win; q++) {
			// ---- accumulation window q: r bins [ra, rb] counted from the top -----------------------------------------------
			const int rb = P.n_r - 1 - q * w_r, ra = (rb - w_r + 1 > 0) ? rb - w_r + 1 : 0;
			const double win_hi = P.r2_thr[rb + 1];
			cx.ra = ra;
			cx.rb = rb;
			cx.lo = P.r2_thr[ra];
			cx.hi = win_hi;
			cx.thr = (ra + 1 <= rb) ? P.r2_thr[ra + 1] : INFINITY;
			cx.hi_lane = dead ? -1.0 : win_hi;  // dead lanes never pass the range test
			const double reach_q = sqrt(win_hi) * (1.0 + 1e-9);

			// ---- keep only the neighbour columns the warp's bounding box can reach in this window (windows shrink, so the
			// list is compacted in place) ------------------------------------------------------------------------------------
			{
				int kept = 0;
				for (int base = 0; base < nn; base += 32) {
					const int c_ = (base + lane < nn) ? nlist[noff + base + lane] : -1;
					bool keep = false;
					if (c_ >= 0) {
						const ColInfo ci = a.colinfo[c_];
						const int cu_ = axis_code(bu0, bu1, ci.umin, ci.umax), cv_ = axis_code(bv0, bv1, ci.vmin, ci.vmax);
						// box to box under the warp's image of the column (straddling: always keep)
						auto box_gap = [&](double b0, double b1, double cmin, double cmax, int code) -> double {
							if (code == 3) return 0.0;
							const double sh = code == 1 ? L : (code == 2 ? -L : 0.0);
							return fmax(0.0, fmax((cmin + sh) - b1, b0 - (cmax + sh)));
						};
						const double gu_ = box_gap(bu0, bu1, ci.umin, ci.umax, cu_), gv_ = box_gap(bv0, bv1, ci.vmin, ci.vmax, cv_);
						keep = (gu_ * gu_ + gv_ * gv_) * (1.0 - 1e-9) < win_hi;
					}
					const unsigned m = __ballot_sync(0xffffffffu, keep);
					__syncwarp();
					if (keep) nlist[kept + __popc(m & ((1u << lane) - 1u))] = c_;
					kept += __popc(m);
				}
				__syncwarp();
				nn = kept;
				noff = 0;  // the compacted list starts at the front
			}

			// line-of-sight regions the warp can reach at all in this window
			int lr_first = 0, lr_count = n_lr;
			if (n_lr > 1) {
				const double wl0 = sl0 - reach_q - 2.0 * eps_l, wl1 = sl1 + reach_q + 2.0 * eps_l;
				if (wl1 - wl0 < L) {
					const int c0 = (int)floor(wl0 * P.inv_cl), c1 = (int)floor(wl1 * P.inv_cl);  // may lie outside [0, nz)
					const int r0 = (int)floor((double)c0 / (double)lr_cells), r1 = (int)floor((double)c1 / (double)lr_cells);
					if (r1 - r0 + 1 < n_lr) {
						lr_first = periodic ? ((r0 % n_lr) + n_lr) % n_lr : (r0 < 0 ? 0 : r0);
						lr_count = periodic ? r1 - r0 + 1 : ((r1 >= n_lr ? n_lr - 1 : r1) - lr_first + 1);
					}
				}
			}

			for (int lri = 0; lri < lr_count; lri++) {
				int g_r = lr_first + lri;
				if (g_r >= n_lr) g_r -= n_lr;
				const int R0 = g_r * lr_cells, R1 = (n_lr > 1) ? R0 + lr_cells - 1 : nz - 1;
				for (int base = 0; base < nn; base += 32) {
					// ---- ROUND: 32 neighbour columns, one per lane --------------------------------------------------------
					const int r_c = (base + lane < nn) ? nlist[base + lane] : -1;
					ColInfo ci;
					ci.umin = ci.vmin = INFINITY;
					ci.umax = ci.vmax = -INFINITY;
					if (r_c >= 0) ci = a.colinfo[r_c];
					const int cu_ = axis_code(bu0, bu1, ci.umin, ci.umax), cv_ = axis_code(bv0, bv1, ci.vmin, ci.vmax);
					// slabs within sqrt(r_hi^2 - d_uv^2) of some shape of the warp: every shape (broadcast by shuffles)
					// against this lane's column, in single precision rounded outwards (a superset is all that is needed)
					float lmin_f = INFINITY, lmax_f = -INFINITY;
					for (int i = 0; i < np; i++) {
						const double xu = __shfl_sync(0xffffffffu, p.u, i), xv = __shfl_sync(0xffffffffu, p.v, i);
						const double xl = __shfl_sync(0xffffffffu, p.l, i);
						const double gu = gap(xu, ci.umin, ci.umax, cu_), gv = gap(xv, ci.vmin, ci.vmax, cv_);
						const double g2 = (gu * gu + gv * gv) * (1.0 - 1e-9);
						if (g2 < win_hi) {
							const double dl = (double)(__fsqrt_ru(__double2float_ru(win_hi - g2)) * 1.000001f) + eps_l;
							lmin_f = fminf(lmin_f, __double2float_rd(xl - dl));
							lmax_f = fmaxf(lmax_f, __double2float_ru(xl + dl));
						}
					}
					int r_sA = 0, r_eA = 0, r_sB = 0, r_eB = 0, r_lab = -2, r_codes = 0, r_slA = 0, r_slB = 0;
					if (r_c >= 0 && lmin_f <= lmax_f) {
						const double lmin = (double)lmin_f, lmax = (double)lmax_f;
						int sa0, sa1, sb0 = 0, sb1 = -1;
						bool none = false;
						if (!periodic) {
							sa0 = cell_index(lmin, P.inv_cl, nz);
							sa1 = cell_index(lmax, P.inv_cl, nz);
							none = (lmax < 0.0 || lmin >= L);
						} else if (!(lmax - lmin < L)) {
							sa0 = 0;
							sa1 = nz - 1;
						} else {
							bool wrapped = false;
							double x0 = lmin, x1 = lmax;
							if (x0 < 0.0) {
								x0 += L;
								wrapped = true;
							}
							if (x1 >= L) {
								x1 -= L;
								wrapped = true;
							}
							const int ca = cell_index(x0, P.inv_cl, nz), cb_ = cell_index(x1, P.inv_cl, nz);
							if (!wrapped) {
								sa0 = ca;
								sa1 = cb_;
							} else if (cb_ >= ca - 1) {
								sa0 = 0;
								sa1 = nz - 1;
							} else {
								sa0 = ca;
								sa1 = nz - 1;
								sb0 = 0;
								sb1 = cb_;
							}
						}
						// intersect with the current line-of-sight region
						sa0 = sa0 > R0 ? sa0 : R0;
						sa1 = sa1 < R1 ? sa1 : R1;
						sb0 = sb0 > R0 ? sb0 : R0;
						sb1 = sb1 < R1 ? sb1 : R1;
						const long long cb0 = (long long)r_c * nz;
						int clA = 0, clB = 0;
						if (!none && sa0 <= sa1) {
							r_sA = (int)a.cell_start[cb0 + sa0];
							r_eA = (int)a.cell_start[cb0 + sa1 + 1];
							r_slA = sa0 | (sa1 << 16);
							clA = axis_code(sl0, sl1, a.slab_lo[sa0], a.slab_hi[sa1]);
						}
						if (!none && sb0 <= sb1) {
							r_sB = (int)a.cell_start[cb0 + sb0];
							r_eB = (int)a.cell_start[cb0 + sb1 + 1];
							r_slB = sb0 | (sb1 << 16);
							clB = axis_code(sl0, sl1, a.slab_lo[sb0], a.slab_hi[sb1]);
						}
						r_codes = cu_ | (cv_ << 2) | (clA << 4) | (clB << 6);
						if (r_eA > r_sA || r_eB > r_sB) r_lab = a.colreg[(long long)r_c * n_lr + g_r];
					}
					const unsigned m_simple = __ballot_sync(0xffffffffu, r_lab >= 0);
					if (m_simple) process_round<UNITW, LOS2, SIG, SYM>(&cx, r_sA, r_eA, r_sB, r_eB, r_lab, r_codes, m_simple);
					// ---- column-regions holding several labels (cells cut by a jackknife face: unaligned grids only) ------------
					unsigned m_cplx = __ballot_sync(0xffffffffu, r_lab == -1);
					while (m_cplx) {
						const int e = __ffs(m_cplx) - 1;
						m_cplx &= m_cplx - 1u;
						const long long cb0 = (long long)__shfl_sync(0xffffffffu, r_c, e) * nz;
						const int codes_e = __shfl_sync(0xffffffffu, r_codes, e);
						const int slA = __shfl_sync(0xffffffffu, r_slA, e), slB = __shfl_sync(0xffffffffu, r_slB, e);
						const int okA = __shfl_sync(0xffffffffu, (int)(r_eA > r_sA), e), okB = __shfl_sync(0xffffffffu, (int)(r_eB > r_sB), e);
						for (int piece = 0; piece < 2; piece++) {
							if (!(piece ? okB : okA)) continue;
							const int s0_ = (piece ? slB : slA) & 0xffff, s1_ = (piece ? slB : slA) >> 16;
							const int pc = (codes_e & 15) | (((codes_e >> (4 + 2 * piece)) & 3) << 4);
							for (int sb = s0_; sb <= s1_; sb += 32) {  // 32 cells at a time, one per lane
								int c_s = 0, c_e = 0, c_lab = -2, c_nlab = 0;
								if (sb + lane <= s1_) {
									c_s = (int)a.cell_start[cb0 + sb + lane];
									c_e = (int)a.cell_start[cb0 + sb + lane + 1];
									const CellInfo *cinf = a.cinfo + cb0 + sb + lane;
									c_nlab = (c_e > c_s) ? cinf->nlab : 0;
									c_lab = (c_nlab == 1) ? cinf->label : -2;
								}
								const unsigned m1 = __ballot_sync(0xffffffffu, c_nlab == 1);
								if (m1) process_round<UNITW, LOS2, SIG, SYM>(&cx, c_s, c_e, 0, 0, c_lab, pc, m1);
								unsigned mm = __ballot_sync(0xffffffffu, c_nlab > 1);
								while (mm) {  // a cell with several labels: one label run at a time
									const int f = __ffs(mm) - 1;
									mm &= mm - 1u;
									int pos = __shfl_sync(0xffffffffu, c_s, f);
									const int end = __shfl_sync(0xffffffffu, c_e, f);
									while (pos < end) {
										const int lb = a.cand_jk[pos];
										int qq = pos + 1;
										while (qq < end && a.cand_jk[qq] == lb) qq++;
										process_round<UNITW, LOS2, SIG, SYM>(&cx, pos, qq, 0, 0, lb, pc, 1u);  // lane 0 carries the run
										pos = qq;
									}
								}
							}
						}
					}
				}
			}
			// ---- end of the window: flush what is left in the private slots --------------------------------------------------
			if (cx.cur_label >= 0) {
				PrivAcc acc;
				acc.a2 = cx.a2;
				acc.aw = cx.aw;
				acc.ac = cx.ac;
				acc.av = cx.av;
				if (SYM)
					cx.binned += flush_slots_rmu_sym<UNITW>(cx.fc, acc, cx.ar_off, p.jk, dead, cx.pe, p.w, ra, rb, n_mu, ns, cx.cur_label);
				else
					cx.binned += flush_slots_rmu<UNITW, SIG>(cx.fc, acc, p.jk, dead, cx.pe, p.w, ra, rb, n_mu, ns, cx.cur_label);
				cx.cur_label = -1;
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

template <bool UNITW, bool LOS2, bool SIG, bool SYM = false>
inline int launch_rmu_variant(const TiledArgs &a, int n_ctas, size_t smem, cudaStream_t st) {
	MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rmu<UNITW, LOS2, SIG, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_tiled_rmu<UNITW, LOS2, SIG, SYM><<<n_ctas, TP, smem, st>>>(a);
	return (int)cudaGetLastError();
}

inline int launch_rmu_sym(const TiledArgs &a, bool unit_w, bool los2, int n_ctas, size_t smem, cudaStream_t st) {
	if (unit_w)
		return los2 ? launch_rmu_variant<true, true, false, true>(a, n_ctas, smem, st)
					: launch_rmu_variant<true, false, false, true>(a, n_ctas, smem, st);
	return los2 ? launch_rmu_variant<false, true, false, true>(a, n_ctas, smem, st)
				: launch_rmu_variant<false, false, false, true>(a, n_ctas, smem, st);
}

template <bool SIG>
inline int launch_rmu_sig(const TiledArgs &a, bool unit_w, bool los2, int n_ctas, size_t smem, cudaStream_t st) {
	if (unit_w) return los2 ? launch_rmu_variant<true, true, SIG>(a, n_ctas, smem, st) : launch_rmu_variant<true, false, SIG>(a, n_ctas, smem, st);
	return los2 ? launch_rmu_variant<false, true, SIG>(a, n_ctas, smem, st) : launch_rmu_variant<false, false, SIG>(a, n_ctas, smem, st);
}

inline int launch_rmu(const TiledArgs &a, bool unit_w, bool los2, bool sig, int n_ctas, size_t smem, cudaStream_t st) {
	return sig ? launch_rmu_sig<true>(a, unit_w, los2, n_ctas, smem, st) : launch_rmu_sig<false>(a, unit_w, los2, n_ctas, smem, st);
}

}  // namespace mia

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

inline size_t tiled_rmu_smem_bytes(bool unit_w) {
	const size_t fixed = sizeof(Cand) * TW * STAGES * CH_RMU + sizeof(int) * TW * MAX_NEIGH + 256 + 768;
	const size_t per_slot = (size_t)TP * (8 + 8 + 4 + (unit_w ? 0 : 8));
	return fixed + per_slot * NS_RMU;
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

// The approximate per-pair quantities of the fast loop, in ONE place: the slow path must reproduce the fast loop's
// "suspect" decisions bit for bit.
struct RmuApprox {
	double gp, gc;  // cos 2phi, sin 2phi
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
	int idx = (int)((thi & 0xFFFFFu) >> 8) - 2049;
	idx = idx < 0 ? 0 : idx;
	r.idx = idx > n_mu - 1 ? n_mu - 1 : idx;
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
	r.susp = susp_mu || (r.gp >= 1.0 - 1e-11);
	return r;
}

// Window of r bins [ra, ra + W_R): limits and interior threshold as bit patterns (s >= 0: integer order == double order)
struct RmuWindow {
	long long lo_b, hi_b, thr_b, cut_b;
	int ra;
};

// One staged chunk against this thread's shape galaxy.  VAR 0: no periodic image; 1: lane-constant image shifts
// (su, sv, sl); 2: every separation wrapped per pair (a chunk straddles +-L/2 for some lane: tiny boxes only).
template <bool UNITW, bool LOS2, int VAR>
__device__ __forceinline__ bool pair_loop_rmu(uint32_t cb, int n, double L, double halfL, double pu, double pv, double pl,
											  double a0, double a1, double su, double sv, double sl, const RmuWindow &rw,
											  long long hi_lane_b, double hn, double tbias, int n_mu, const PrivAcc &acc) {
	auto wrap = [&](double d) {
		const double c = __hiloint2double(__double2hiint(L) | (__double2hiint(d) & 0x80000000), __double2loint(L));
		return (fabs(d) > halfL) ? __dsub_rn(d, c) : d;  // c = copysign(L, d): measure_m_box_jk.py:419-421
	};
	bool lane_susp = false;
	double cu, cv, cl, cw, mu_, mv_, ml_, mw_;
	lds_v2(cu, cv, cb);
	lds_v2(cl, cw, cb + 16);
	{
		const uint32_t a1_ = cb + (uint32_t)((1 < n) ? 1 : 0) * (uint32_t)sizeof(Cand);
		lds_v2(mu_, mv_, a1_);
		lds_v2(ml_, mw_, a1_ + 16);
	}
	MIA_UNROLL_PRAGMA(MIA_UNROLL)
	for (int j = 0; j < n; j++) {
		const uint32_t na = cb + (uint32_t)((j + 2 < n) ? (j + 2) : (n - 1)) * (uint32_t)sizeof(Cand);
		double nu, nv, nl, nw;
		lds_v2(nu, nv, na);
		lds_v2(nl, nw, na + 16);
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
		const long long sb = __double_as_longlong(s);
		bool ok = (sb >= rw.lo_b) && (sb < hi_lane_b) && (__double_as_longlong(rp2) > rw.cut_b);
		const RmuApprox ap = rmu_approx(du, dv, dz, rp2, s, a0, a1, hn, tbias, n_mu);
		const int slot = ap.idx + ((sb >= rw.thr_b) ? n_mu : 0);
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double s0, s1, sw = 0.0;
		lds_v2(s0, s1, acc.a2 + so * 16u);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) sw = lds_f64(acc.aw + so * 8u);
		lane_susp = lane_susp || (ok && ap.susp);
		ok = ok && !ap.susp;
		double gp = ap.gp, gc = ap.gc;
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64_if(ok, acc.aw + so * 8u, sw + cw);
		}
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
template <bool UNITW, bool LOS2>
__device__ __noinline__ void slow_pairs_rmu(bool lane_susp, uint32_t cb, int n, int periodic, double L, double halfL,
											double pu, double pv, double pl, double a0, double a1, const RmuWindow rw,
											long long hi_lane_b, double hn, double tbias, int n_mu, const double *thr2,
											PrivAcc acc, unsigned long long &nan_pairs) {
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
		lds_v2(cu, cv, cb + (uint32_t)j * (uint32_t)sizeof(Cand));
		lds_v2(cl, cw, cb + (uint32_t)j * (uint32_t)sizeof(Cand) + 16);
		const double du = sep(pu, cu), dv = sep(pv, cv), dz = sep(pl, cl);
		const double uu = __dmul_rn(du, du), vv = __dmul_rn(dv, dv), ll = __dmul_rn(dz, dz);
		const double rp2 = __dadd_rn(uu, vv);
		const double s = LOS2 ? __dadd_rn(rp2, ll) : __dadd_rn(__dadd_rn(uu, ll), vv);
		const long long sb = __double_as_longlong(s);
		if (!((sb >= rw.lo_b) && (sb < hi_lane_b) && (__double_as_longlong(rp2) > rw.cut_b))) continue;
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
		const int slot = idx + ((sb >= rw.thr_b) ? n_mu : 0);
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double s0, s1;
		lds_v2(s0, s1, acc.a2 + so * 16u);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) {
			gp *= cw;
			gc *= cw;
			sts_f64(acc.aw + so * 8u, lds_f64(acc.aw + so * 8u) + cw);
		}
		sts_v2(acc.a2 + so * 16u, s0 + gp, s1 + gc);
		sts_u32(acc.ac + so * 4u, c0 + 1u);
	}
}

// Flush: fixed-order warp reduction of the private slots into this warp's accumulator copy in HBM.  All lanes share the
// slot -> bin map (slot = r_offset * n_mu + mu bin); lanes are grouped by the jackknife label of their shape galaxy.
template <bool UNITW>
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
		double tot_sp = 0.0, tot_sc = 0.0, tot_dw = 0.0;
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
			if (lane == sl) {
				tot_cnt = csum;
				tot_sp = xs;
				tot_sc = ys;
				tot_dw = zs;
			}
		}
		if (lane < ns && tot_cnt) {
			const int roff = lane / n_mu, b2 = lane - roff * n_mu;
			const int rbin = ra + roff;
			if (rbin > rb) {
				atomicExch(&fc.flags[1], 1);
			} else {
				const size_t bin = (size_t)rbin * fc.n_2 + b2;
				const size_t ia = (size_t)k * fc.nb + bin;
				fc.pcnt[ia] += tot_cnt;
				fc.pddw[ia] += tot_dw;
				fc.psp[ia] += tot_sp;
				fc.psc[ia] += tot_sc;
				if (fc.num_jk > 0 && jkD != k) {
					const size_t ib = (size_t)(fc.J + jkD) * fc.nb + bin;
					fc.pcnt[ib] += tot_cnt;
					fc.pddw[ib] += tot_dw;
					fc.psp[ib] += tot_sp;
				}
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

struct ChunkRmu {
	long long start;
	int n, label;
	int xy;        // projected axes: 0 no image, 1 lane-constant image shifts, 2 some lane straddles +-L/2
	int cl0, cl1;  // first / last non-empty slab of the chunk's run (bounds of its line-of-sight coordinates)
};

template <bool UNITW, bool LOS2>
__global__ void __launch_bounds__(TP, MIA_MIN_CTAS) k_tiled_rmu(const TiledArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	const DevParams &P = a.P;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb = P.n_r * P.n_2;
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	const int periodic = P.periodic;
	const double L = P.L, halfL = P.halfL;
	const int n_mu = P.n_2, w_r = a.w_r, ns = w_r * n_mu;
	const int nz = a.nz;
	const double hn = 0.5 * (double)n_mu, tbias = hn + 6145.0;

	// ---- shared memory carve-up ------------------------------------------------------------------------------------
	Cand *ring = reinterpret_cast<Cand *>(smem);  // [warp][stage][CH_RMU]
	int *nlist_all = reinterpret_cast<int *>(smem + sizeof(Cand) * TW * STAGES * CH_RMU);
	uint64_t *full = reinterpret_cast<uint64_t *>(nlist_all + TW * MAX_NEIGH);  // [warp][stage]
	double *thr2_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(full) + 256);
	unsigned char *accbase = reinterpret_cast<unsigned char *>(thr2_s) + 768;
	const uint32_t acc_u32 = smem_u32(accbase);
	Cand *my_ring = ring + (size_t)warp * STAGES * CH_RMU;
	uint64_t *my_full = full + warp * STAGES;
	int *nlist = nlist_all + warp * MAX_NEIGH;
	const uint32_t my_ring_u32 = smem_u32(my_ring);
	PrivAcc acc;
	acc.a2 = acc_u32 + (uint32_t)tid * 16u;
	acc.aw = acc_u32 + (uint32_t)NS_RMU * TP * 16u + (uint32_t)tid * 8u;
	acc.ac = acc_u32 + (uint32_t)NS_RMU * TP * (UNITW ? 16u : 24u) + (uint32_t)tid * 4u;

	if (tid == 0) {
		for (int s = 0; s < TW * STAGES; s++) mbar_init(&full[s], 1);
		mbar_fence_init();
		if (blockIdx.x == 0) a.A.stats[6] = (unsigned long long)a.n_tasks[0];
	}
	for (int e = tid; e <= P.n_2; e += blockDim.x) thr2_s[e] = P.thr2[e];
#pragma unroll 1
	for (int s = 0; s < NS_RMU; s++) {
		sts_v2(acc.a2 + (uint32_t)s * TP * 16u, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)s * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)s * TP * 4u, 0u);
	}
	__syncthreads();  // the only CTA-wide synchronisation

	// ---- this warp's share of the tasks (same rule as the (r_p, Pi) kernel) -------------------------------------------
	int task0 = 0, task1 = 0;
	{
		const int nt = a.n_tasks[0];
		if (nt > 0) {
			const double total2 = 2.0 * (double)a.task_cum[nt - 1];
			const int RG = a.shard_count * a.n_workers;
			const int mine = a.shard_index * a.n_workers + (int)blockIdx.x * TW + warp;
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

	const size_t part = (size_t)(blockIdx.x * TW + warp) * (size_t)a.A.rows * nb;
	FlushCtx fc;
	fc.pcnt = a.A.cnt + part;
	fc.pddw = a.A.ddw + part;
	fc.psp = a.A.sp + part;
	fc.psc = a.A.sc + part;
	fc.flags = a.flags;
	fc.n_2 = P.n_2;
	fc.nb = nb;
	fc.J = J;
	fc.num_jk = P.num_jk;

	uint32_t phase0 = 0u, phase1 = 0u;
	int st_issue = 0;
	unsigned long long tested = 0, binned = 0, nan_pairs = 0;
	const double TN = P.r2_thr[P.n_r];
	const double cs = L / P.ncu, reach = sqrt(TN) * (1.0 + 1e-6);
	const int n_win = (P.n_r + w_r - 1) / w_r;
	// line-of-sight regions: when the slabs are aligned with the jackknife sub-boxes, candidates are visited region by
	// region so that the label of consecutive chunks changes as rarely as possible
	const int n_lr = (a.n_side > 1 && nz % a.n_side == 0) ? a.n_side : 1;
	const int lr_cells = nz / n_lr;
	const double eps_l = 1e-9 * L;

	for (int task = task0; task < task1; task++) {
		const int col = a.task_col[task];
		const int np = a.task_n[task];
		const bool dead = lane >= np;
		Prim p;
		if (!dead) {
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
		const int nn = build_neighbour_list(nlist, col, P.ncu, P.ncv, P.ku, P.kv, periodic, a.n_side, cs, reach);

		for (int q = 0; q < n_win; q++) {
			// ---- accumulation window q: r bins [ra, rb] counted from the top -----------------------------------------------
			const int rb = P.n_r - 1 - q * w_r, ra = (rb - w_r + 1 > 0) ? rb - w_r + 1 : 0;
			RmuWindow rw;
			rw.ra = ra;
			rw.lo_b = __double_as_longlong(P.r2_thr[ra]);
			rw.hi_b = __double_as_longlong(P.r2_thr[rb + 1]);
			rw.thr_b = (ra + 1 <= rb) ? __double_as_longlong(P.r2_thr[ra + 1]) : 0x7ff0000000000000ll;
			rw.cut_b = __double_as_longlong(P.rp2_cut);
			const double win_hi = P.r2_thr[rb + 1];
			const long long hi_lane_b = dead ? 0ll : rw.hi_b;  // dead lanes never pass the range test
			const double reach_q = sqrt(win_hi) * (1.0 + 1e-9);

			auto flush = [&](int jkD) {
				binned += flush_slots_rmu<UNITW>(fc, acc, p.jk, dead, pe, p.w, ra, rb, n_mu, ns, jkD);
			};

			// line-of-sight regions the warp can reach at all in this window
			int lr_first = 0, lr_count = n_lr;
			if (n_lr > 1) {
				const double wl0 = warp_min_f64(dead ? INFINITY : p.l) - reach_q - eps_l;
				const double wl1 = warp_max_f64(dead ? -INFINITY : p.l) + reach_q + eps_l;
				if (wl1 - wl0 < L) {
					const int c0 = (int)floor(wl0 * P.inv_cl), c1 = (int)floor(wl1 * P.inv_cl);  // may lie outside [0, nz)
					const int r0 = (int)floor((double)c0 / (double)lr_cells), r1 = (int)floor((double)c1 / (double)lr_cells);
					if (r1 - r0 + 1 < n_lr) {
						lr_first = periodic ? ((r0 % n_lr) + n_lr) % n_lr : (r0 < 0 ? 0 : r0);
						lr_count = periodic ? r1 - r0 + 1 : ((r1 >= n_lr ? n_lr - 1 : r1) - lr_first + 1);
					}
				}
			}

			// ---- chunk generator: region -> neighbour column -> segment of slabs -> batch of 32 cells -> label run -> chunk
			int g_lri = 0;                        // regions done
			int g_R0 = 0, g_R1 = -1;              // slab range of the current region
			int g_ci = nn;                        // next neighbour column (nn: open the next region first)
			int g_xy = 0;
			double g_su = 0.0, g_sv = 0.0;
			long long g_colbase = 0;
			int g_seg_next = 0, g_seg_last = -1;  // current segment of slabs (inclusive)
			int g_seg2_lo = 0, g_seg2_hi = -1;    // pending second segment
			long long d_start = 0, d_end = 0;     // per lane: candidate range of one cell of the batch
			int d_label = -1, d_nlab = 0;
			unsigned g_runs = 0u, g_ne = 0u;
			int g_batch_c0 = 0;
			long long g_pos = 0, g_run_end = 0, g_cell_end = 0;
			int g_label = -1, g_cl0 = 0, g_cl1 = 0;

			auto next_chunk = [&](ChunkRmu &c) -> bool {
				for (;;) {
					if (g_pos < g_run_end) {
						const int rest = (int)((g_run_end - g_pos < 1000000) ? (g_run_end - g_pos) : 1000000);
						const int nch = (rest + CH_RMU - 1) / CH_RMU;
						c.start = g_pos;
						c.n = (rest + nch - 1) / nch;
						c.label = g_label;
						c.xy = g_xy;
						c.cl0 = g_cl0;
						c.cl1 = g_cl1;
						g_pos += c.n;
						return true;
					}
					if (g_pos < g_cell_end) {  // next label run of a cell that holds several labels
						g_label = a.cand_jk[g_pos];
						long long qq = g_pos + 1;
						while (qq < g_cell_end && a.cand_jk[qq] == g_label) qq++;
						g_run_end = qq;
						continue;
					}
					if (g_runs) {  // next run of the batch
						const int e = __ffs(g_runs) - 1;
						g_runs &= g_runs - 1u;
						const int nx = g_runs ? __ffs(g_runs) - 1 : 32;
						const unsigned below = g_ne & (nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u));
						const int last = 31 - __clz(below);  // e itself is non-empty and below nx
						g_pos = __shfl_sync(0xffffffffu, d_start, e);
						g_cell_end = __shfl_sync(0xffffffffu, d_end, last);
						g_label = __shfl_sync(0xffffffffu, d_label, e);
						const int nlab = __shfl_sync(0xffffffffu, d_nlab, e);
						g_cl0 = g_batch_c0 + e;
						g_cl1 = g_batch_c0 + last;
						g_run_end = g_cell_end;
						if (nlab > 1) g_run_end = g_pos;  // a cell with several labels: split it into its label runs (above)
						continue;
					}
					if (g_seg_next <= g_seg_last) {  // next batch of up to 32 cells of the segment
						const int nbatch = (g_seg_last - g_seg_next + 1 < 32) ? (g_seg_last - g_seg_next + 1) : 32;
						d_start = d_end = 0;
						d_label = -1;
						d_nlab = 0;
						if (lane < nbatch) {
							const long long cc = g_colbase + g_seg_next + lane;
							d_start = a.cell_start[cc];
							d_end = a.cell_start[cc + 1];
							const CellInfo *ci = a.cinfo + cc;
							d_label = ci->label;
							d_nlab = ci->nlab;
						}
						const unsigned ne = __ballot_sync(0xffffffffu, d_end > d_start);
						const unsigned below = ne & ((1u << lane) - 1u);
						const int prev = below ? 31 - __clz(below) : 0;
						const int prev_label = __shfl_sync(0xffffffffu, d_label, prev);
						const int prev_nlab = __shfl_sync(0xffffffffu, d_nlab, prev);
						const bool boundary = (d_end > d_start) &&
											  (!below || d_label != prev_label || d_nlab > 1 || prev_nlab > 1);
						g_runs = __ballot_sync(0xffffffffu, boundary);
						g_ne = ne;
						g_batch_c0 = g_seg_next;
						g_seg_next += nbatch;
						continue;
					}
					if (g_seg2_lo <= g_seg2_hi) {
						g_seg_next = g_seg2_lo;
						g_seg_last = g_seg2_hi;
						g_seg2_hi = -1;
						g_seg2_lo = 0;
						continue;
					}
					if (g_ci >= nn) {  // next line-of-sight region
						if (g_lri >= lr_count) return false;
						int r = lr_first + g_lri;
						if (r >= n_lr) r -= n_lr;
						g_lri++;
						g_R0 = r * lr_cells;
						g_R1 = (n_lr > 1) ? g_R0 + lr_cells - 1 : nz - 1;
						g_ci = 0;
						continue;
					}
					// ---- open the next neighbour column: per-lane image shifts, culling, range of slabs ---------------------
					const int ncol_ = nlist[g_ci++];
					const ColInfo ci = a.colinfo[ncol_];
					double ulo = __dsub_rn(p.u, ci.umax), uhi = __dsub_rn(p.u, ci.umin);
					double vlo = __dsub_rn(p.v, ci.vmax), vhi = __dsub_rn(p.v, ci.vmin);
					bool strad = false;
					double su = 0.0, sv = 0.0;
					if (periodic) {
						if (!(ulo >= -halfL && uhi <= halfL)) {
							if (ulo > halfL) {  // every pair of this lane with the column wraps down
								ulo = __dsub_rn(ulo, L);
								uhi = __dsub_rn(uhi, L);
								su = -L;
							} else if (uhi < -halfL) {
								ulo = __dadd_rn(ulo, L);
								uhi = __dadd_rn(uhi, L);
								su = L;
							} else {
								strad = true;
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
								strad = true;
							}
						}
					}
					const double mu = ulo > 0.0 ? ulo : (uhi < 0.0 ? -uhi : 0.0);
					const double mv = vlo > 0.0 ? vlo : (vhi < 0.0 ? -vhi : 0.0);
					const double d2 = strad ? 0.0 : __dadd_rn(__dmul_rn(mu, mu), __dmul_rn(mv, mv));
					const bool need = !dead && (d2 < win_hi);
					if (!__any_sync(0xffffffffu, need)) continue;
					// slabs within sqrt(r_hi^2 - d_uv^2) of some shape of the warp
					double lmin = INFINITY, lmax = -INFINITY;
					if (need) {
						const double dl = sqrt(win_hi - d2) * (1.0 + 1e-9) + eps_l;
						lmin = p.l - dl;
						lmax = p.l + dl;
					}
					lmin = warp_min_f64(lmin);
					lmax = warp_max_f64(lmax);
					int sa0, sa1, sb0 = 0, sb1 = -1;
					if (!periodic) {
						sa0 = cell_index(lmin, P.inv_cl, nz);
						sa1 = cell_index(lmax, P.inv_cl, nz);
						if (lmax < 0.0 || lmin >= L) continue;
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
					sa0 = sa0 > g_R0 ? sa0 : g_R0;
					sa1 = sa1 < g_R1 ? sa1 : g_R1;
					sb0 = sb0 > g_R0 ? sb0 : g_R0;
					sb1 = sb1 < g_R1 ? sb1 : g_R1;
					if (sa0 > sa1 && sb0 > sb1) continue;
					g_xy = __any_sync(0xffffffffu, !dead && strad) ? 2
						   : (__any_sync(0xffffffffu, !dead && (su != 0.0 || sv != 0.0)) ? 1 : 0);
					g_su = su;
					g_sv = sv;
					g_colbase = (long long)ncol_ * nz;
					if (sa0 <= sa1) {
						g_seg_next = sa0;
						g_seg_last = sa1;
						g_seg2_lo = sb0;
						g_seg2_hi = sb1;
					} else {
						g_seg_next = sb0;
						g_seg_last = sb1;
						g_seg2_lo = 0;
						g_seg2_hi = -1;
					}
				}
			};

			// ---- software pipeline: issue the bulk copy of chunk k+1, then work on chunk k ----------------------------------
			ChunkRmu pend, nxt;
			double pend_su = 0.0, pend_sv = 0.0, nxt_su = 0.0, nxt_sv = 0.0;
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
					bulk_load(my_ring + (size_t)st_issue * CH_RMU, a.cand + nxt.start, bytes, &my_full[st_issue]);
				}
				const int lab = (pend.n > 0) ? pend.label : -2;
				if (lab != cur_label) {
					if (cur_label >= 0) flush(cur_label);
					cur_label = lab;
				}
				if (pend.n > 0) {
					// line-of-sight image of this lane for the chunk (its candidates lie in [zlo, zhi])
					const double zlo = a.slab_lo[pend.cl0], zhi = a.slab_hi[pend.cl1];
					double sl = 0.0;
					bool zst = false;
					if (periodic) {
						const double lo = __dsub_rn(p.l, zhi), hi = __dsub_rn(p.l, zlo);
						if (!(lo >= -halfL && hi <= halfL)) {
							if (lo > halfL) sl = -L;
							else if (hi < -halfL) sl = L;
							else zst = true;
						}
					}
					const bool gen = (pend.xy == 2) || __any_sync(0xffffffffu, !dead && zst);
					const bool shifted =
						(pend.xy == 1) || __any_sync(0xffffffffu, !dead && sl != 0.0);
					if (pend_st == 0) {
						mbar_wait(&my_full[0], phase0);
						phase0 ^= 1u;
					} else {
						mbar_wait(&my_full[1], phase1);
						phase1 ^= 1u;
					}
					if (!dead) tested += (unsigned long long)pend.n;
					const uint32_t cb = my_ring_u32 + (uint32_t)pend_st * (uint32_t)(CH_RMU * sizeof(Cand));
					bool susp;
					if (gen)
						susp = pair_loop_rmu<UNITW, LOS2, 2>(cb, pend.n, L, halfL, p.u, p.v, p.l, p.a0, p.a1, 0.0, 0.0, 0.0, rw,
															 hi_lane_b, hn, tbias, n_mu, acc);
					else if (shifted)
						susp = pair_loop_rmu<UNITW, LOS2, 1>(cb, pend.n, L, halfL, p.u, p.v, p.l, p.a0, p.a1, pend_su, pend_sv, sl,
															 rw, hi_lane_b, hn, tbias, n_mu, acc);
					else
						susp = pair_loop_rmu<UNITW, LOS2, 0>(cb, pend.n, L, halfL, p.u, p.v, p.l, p.a0, p.a1, 0.0, 0.0, 0.0, rw,
															 hi_lane_b, hn, tbias, n_mu, acc);
					if (__any_sync(0xffffffffu, susp))
						slow_pairs_rmu<UNITW, LOS2>(susp, cb, pend.n, periodic, L, halfL, p.u, p.v, p.l, p.a0, p.a1, rw, hi_lane_b,
													hn, tbias, n_mu, thr2_s, acc, nan_pairs);
					__syncwarp();
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
			if (cur_label >= 0) flush(cur_label);
		}
	}

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

template <bool UNITW, bool LOS2>
inline int launch_rmu_variant(const TiledArgs &a, int n_ctas, size_t smem, cudaStream_t st) {
	MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rmu<UNITW, LOS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_tiled_rmu<UNITW, LOS2><<<n_ctas, TP, smem, st>>>(a);
	return (int)cudaGetLastError();
}

inline int launch_rmu(const TiledArgs &a, bool unit_w, bool los2, int n_ctas, size_t smem, cudaStream_t st) {
	if (unit_w) return los2 ? launch_rmu_variant<true, true>(a, n_ctas, smem, st) : launch_rmu_variant<true, false>(a, n_ctas, smem, st);
	return los2 ? launch_rmu_variant<false, true>(a, n_ctas, smem, st) : launch_rmu_variant<false, false>(a, n_ctas, smem, st);
}

}  // namespace mia

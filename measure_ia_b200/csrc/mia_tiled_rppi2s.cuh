// mia_tiled_rppi2s.cuh -- SYMMETRIC row-streaming (r_p, Pi) pair kernel for AUTO-correlations (sm_100a).
//
// When the position sample and the shape sample are the same catalogue (every BASELINE configuration but cfg5), the ordered
// pairs (shape i, position j) and (shape j, position i) share the separation, r_p^2, the range tests, the r_p bin and
// 1 / r_p^2: sep_ji = -sep_ij EXACTLY in IEEE arithmetic (fl(a - b) = -fl(b - a), and the reference's two conditional
// shifts by +-L commute with negation, measure_w_box_jk.py:401-404).  This kernel therefore visits every UNORDERED pair once
// and accumulates both orderings:
//   * forward  (shape = the lane's galaxy, position = the streamed candidate)  -- as in mia_tiled_rppi2.cuh;
//   * reverse  (shape = the streamed candidate, position = the lane's galaxy): the candidate record carries its own axis
//     and w * e (CandSU / CandSW, 48 / 64 B), Pi -> -Pi (its own pair of Pi slots), rows A[label of the chunk] / B[label of the lane].
// An unordered pair is taken by the galaxy that sees the other one AHEAD along u (wrapped d_u < 0; the rare exact ties
// d_u == 0 are broken lexicographically by the slow path), so a warp only streams its own u rows and the k rows ahead:
// half the rows, half the rounds, the same chunk sizes.
//
// Everything that decides a bin is still the reference's operation sequence compared against the calibrated thresholds; the
// reverse Pi slot is decided by its own comparison (-Pi >= t  <=>  Pi <= -t), and pairs for which forward and reverse
// disagree with the mirrored layout (Pi exactly on a calibrated edge), pairs with d_u == 0, and pairs with |cos| within
// 1e-11 of 1 in either direction go to an exact slow path that adds straight into the warp's accumulator copy, one lane at
// a time (fixed order).  DD stays bit-exact and every sum bit-reproducible.
#pragma once
#include "mia_tiled_rppi2.cuh"

namespace mia {

// Tuning (B200, cfg2, profiles/r02_tuning.md): 2 resident CTAs per SM with 255 registers and no spills in the chunk consumer
// beat 3 CTAs at the 168-register cap (113 vs 120 ms); chunks of 128 (unit weights, 48-byte records) / 96 (64-byte records).
#ifndef MIA_S_CH_U
#define MIA_S_CH_U 128
#endif
#ifndef MIA_S_CH_W
#define MIA_S_CH_W 96
#endif
#ifndef MIA_S_WR
#define MIA_S_WR 5
#endif
#ifndef MIA_S_UNROLL
#define MIA_S_UNROLL 2
#endif
#ifndef MIA_S_MIN_CTAS
#define MIA_S_MIN_CTAS 2
#endif
constexpr int WRS = MIA_S_WR;    // r bins per accumulation window (the same for both variants: it fixes the grouping of the sums)
constexpr int NSLOT_S = 2 * WRS; // private slots per thread: (r bin of the window) x (Pi slot of the forward pair)

// (candidate records CandSU / CandSW: mia_tiled.cuh)
template <bool UNITW>
struct CandRec {
	typedef CandSW type;
	static constexpr int CH = MIA_S_CH_W;  // candidates per staged chunk
};
template <>
struct CandRec<true> {
	typedef CandSU type;
	static constexpr int CH = MIA_S_CH_U;
};
inline int rppi2s_cand_bytes(bool unit_w) { return unit_w ? (int)sizeof(CandSU) : (int)sizeof(CandSW); }

inline size_t tiled_rppi2s_smem_bytes(bool unit_w) {
	const size_t ring = unit_w ? sizeof(CandSU) * CandRec<true>::CH : sizeof(CandSW) * CandRec<false>::CH;
	const size_t fixed = ring * TW * STAGES + 256 + 768;
	const size_t per_slot = (size_t)TP * (16 + 16 + 4 + (unit_w ? 0 : 8));
	return fixed + per_slot * NSLOT_S + (size_t)TP * 16;  // + per-thread flush factors {w e, w}
}

// Can the symmetric kernel take this grid?  The rows a warp streams (its own `ratio` rows and k rows ahead) must be
// unambiguously ahead under the periodic wrap: (k + ratio) cells < L / 2.
inline bool rppi2s_supported(int ncu, int k, int ratio) { return 2 * (k + ratio) < ncu; }

__global__ void k_gather_cands(const double *__restrict__ pos, const double *__restrict__ w, const int32_t *__restrict__ jk,
							   const double *__restrict__ axis, const double *__restrict__ e, const int32_t *__restrict__ idx,
							   int64_t n, int nl0, int nl1, int los, void *__restrict__ out, int32_t *__restrict__ out_jk) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	int64_t s = idx[i];
	CandSW c;
	c.u = pos[3 * s + nl0];
	c.v = pos[3 * s + nl1];
	c.l = pos[3 * s + los];
	c.w = w ? w[s] : 1.0;
	c.a0 = axis[2 * s];
	c.a1 = axis[2 * s + 1];
	c.we = c.w * e[s];
	c.pad = 0.0;
	if (w) {
		reinterpret_cast<CandSW *>(out)[i] = c;
	} else {
		CandSU q;
		q.u = c.u;
		q.v = c.v;
		q.l = c.l;
		q.we = c.we;
		q.a0 = c.a0;
		q.a1 = c.a1;
		reinterpret_cast<CandSU *>(out)[i] = q;
	}
	out_jk[i] = jk ? jk[s] : 0;
}

struct RWindowS {
	double lo, hi;
	double thr[WRS > 1 ? WRS - 1 : 1];
	int ra;
};

// thread-private accumulators (shared-memory byte addresses of slot 0 of this thread):
//   af : [slot][thread] double2 {sum e+, sum ex} of the forward pairs; the reverse pairs' double2 sits AR_OFF bytes further
//   ac : [slot][thread] u32 unordered pairs                                 aw : [slot][thread] sum w_D (weighted variant)
constexpr uint32_t AR_OFF = (uint32_t)NSLOT_S * TP * 16u;
struct PrivAccS {
	uint32_t af, ac, aw;
	uint32_t fac;  // [thread] double2 {w e, w} of the thread's galaxy: the factors the flush applies to the lanes' sums
};

// Loop variants
//   Z_PLAIN : no periodic image in the projected axes, lane-constant image along the line of sight, slab inside the Pi range
//   Z_WRAP  : warp-constant images (su, sv) in the projected axes; line of sight wrapped per pair with the lane's one-sided rule
//             (slab straddling +-L/2); the Pi range is the whole box, so only |Pi| > zb (an exact edge) leaves the mirrored layout
//   Z_GEN   : everything wrapped per pair, Pi range tested for both orderings (Pi range inside the box, tiny boxes)
enum { Z_PLAIN = 0, Z_WRAP = 1, Z_GEN = 2 };

// Per lane and slab: Pi windows of the forward and the reverse pair.
struct ZSym {
	double tf;    // forward:  Pi slot 1 iff dz >= tf
	double trp;   // reverse:  expected (dz > trp) == (dz >= tf); anything else goes to the slow path
	double f_lo, f_hi;  // forward in range iff f_lo <= dz < f_hi          (Z_GEN only)
	double r_lo, r_hi;  // reverse in range iff r_lo <  dz <= r_hi  (i.e. thr2[0] <= -dz < thr2[n])
	double shift;       // lane-constant periodic image along the line of sight
	double wthr, wadd;  // Z_WRAP: a pair takes the image `wadd` iff (dz > wthr) != wflip, else `shift` (see ZWindow)
	double zb;          // Z_WRAP: both orderings are inside the Pi range when |dz| <= zb
	int fb0, fb1;       // forward Pi bin of slot 0 / 1
	int rb0, rb1;       // reverse Pi bin of slot 0 / 1
	int mode;           // Z_* this lane needs for this slab
	bool wflip, dead, err;
};

__device__ __noinline__ ZSym z_window_sym(double pl, double zlo, double zhi, const ZParams P) {
	const ZWindow f = z_window_sep(__dsub_rn(pl, zhi), __dsub_rn(pl, zlo), P);
	const ZWindow r = z_window_sep(__dsub_rn(zlo, pl), __dsub_rn(zhi, pl), P);  // -(pl - c) = c - pl exactly
	ZSym z;
	z.err = f.err || r.err;
	z.dead = f.dead && r.dead;
	z.shift = f.shift;
	z.wthr = f.wthr;
	z.wadd = f.wadd;
	z.wflip = f.wflip;
	z.f_lo = f.t_lo;
	z.f_hi = f.t_hi;
	z.r_lo = -r.t_hi;
	z.r_hi = -r.t_lo;
	z.fb0 = f.b0;
	z.fb1 = f.b1;
	z.zb = INFINITY;
	z.mode = Z_PLAIN;
	if (f.gen || r.gen) {
		// a slab straddling +-L/2 while the Pi range is (to within the calibration of its two outer edges) the whole box:
		// wrapped separations only leave the range on an exact edge
		const double lo = P.thr2[0], hi = P.thr2[P.n_2];
		const double prev_hi = __longlong_as_double(__double_as_longlong(hi) - 1);
		const bool whole_box = P.periodic && lo <= -P.halfL * (1.0 - 1e-12) && hi >= P.halfL * (1.0 - 1e-12);
		const bool straddles = (f.wthr < INFINITY) && (r.wthr < INFINITY);
		if (whole_box && straddles) {
			z.mode = Z_WRAP;
			z.zb = fmin(-lo, prev_hi);
		} else {
			z.mode = Z_GEN;
		}
	}
	const bool f_split = f.t_split < INFINITY, r_split = r.t_split < INFINITY;
	if (f.dead != r.dead || f_split != r_split) {
		// forward and reverse windows of different shape (possible only when a calibrated edge sits exactly on a slab
		// boundary): every pair of this lane with this slab is decided by the slow path
		z.tf = INFINITY;
		z.trp = -INFINITY;
		z.rb0 = z.rb1 = -1;
		z.fb0 = z.fb1 = -1;
		z.mode = Z_GEN;
	} else if (f_split) {
		z.tf = f.t_split;
		z.trp = -r.t_split;  // reverse slot 1 iff -dz >= t  <=>  !(dz > -t)
		z.rb0 = r.b1;        // forward slot 0 (dz below the split) <-> reverse slot 1
		z.rb1 = r.b0;
	} else {
		z.tf = INFINITY;
		z.trp = INFINITY;
		z.rb0 = r.b0;
		z.rb1 = -1;
	}
	return z;
}

// |sin 2phi| at or below this (hi word of the double) sends a pair to the exact path: it covers |cos| within 1e-11 of 1, where
// the reference's NaN rule may apply (and, harmlessly, |cos| ~ 0).  2^-18 = 3.8e-6: about 5e-6 of all pairs.
constexpr int GC_SMALL_HI = 0x3ed00000;

// per-pair geometry shared by pair_loop_sym and slow_pairs_sym: identical code, so that both make identical decisions
struct PairS {
	double r2, inv2, crf, srf, crr, srr;
};
__device__ __forceinline__ void pair_geometry(double du, double dv, double a0, double a1, double b0, double b1, PairS &g) {
	g.r2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));  // measure_w_box_jk.py:407 (before the sqrt)
	// 2 / r2 (hardware seed + one cubic step, see mia_tiled.cuh)
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(g.r2));
	{
		const double e = fma(-g.r2, y, 1.0);
		y = fma(y, fma(e, e, e), y);
	}
	g.inv2 = __hiloint2double(__double2hiint(y) + 0x00100000, __double2loint(y));
	g.crf = fma(du, a0, __dmul_rn(dv, a1));   // forward: the lane's axis
	g.srf = fma(du, a1, -__dmul_rn(dv, a0));
	g.crr = fma(du, b0, __dmul_rn(dv, b1));   // reverse: separation -sep, the candidate's axis: cos = -crr / r_p
	g.srr = fma(du, b1, -__dmul_rn(dv, b0));
}

// One staged chunk against this thread's galaxy, both orderings of every pair with wrapped d_u < 0.
// Returns true when some pair of this lane was left to slow_pairs_sym().
template <bool UNITW, int MODE>
__device__ __forceinline__ bool pair_loop_sym(uint32_t cb, int n, int periodic, double L, double halfL, double pu, double pv,
											  double pl, double a0, double a1, double su, double sv, const RWindowS &rw,
											  double hi_lane, const ZSym &z, const PrivAccS &acc) {
	constexpr uint32_t REC = (uint32_t)sizeof(typename CandRec<UNITW>::type);
	auto wrap = [&](double d) {
		const double sl = __hiloint2double(__double2hiint(L) | (__double2hiint(d) & 0x80000000), __double2loint(L));
		return (fabs(d) > halfL) ? __dsub_rn(d, sl) : d;
	};
	unsigned lane_susp = 0u;
	double cu, cv, cl, we, b0, b1, cw = 1.0, pad_;
	lds_v2(cu, cv, cb);
	lds_v2(cl, we, cb + 16);
	lds_v2(b0, b1, cb + 32);
	if (!UNITW) lds_v2(cw, pad_, cb + 48);
	uint32_t na = cb + REC;
	MIA_UNROLL_PRAGMA(MIA_S_UNROLL)
	for (int j = 0; j < n; j++) {
		// next candidate (one record past the end of the chunk is still inside this CTA's shared memory; never used)
		double nu, nv, nl, nwe, nb0, nb1, nw = 1.0;
		lds_v2(nu, nv, na);
		lds_v2(nl, nwe, na + 16);
		lds_v2(nb0, nb1, na + 32);
		if (!UNITW) lds_v2(nw, pad_, na + 48);
		na += REC;
		double du = __dsub_rn(pu, cu), dv = __dsub_rn(pv, cv), dz = __dsub_rn(pl, cl);  // shape minus position, :401
		if (MODE == Z_GEN) {
			if (periodic) {
				du = wrap(du);
				dv = wrap(dv);
				dz = wrap(dz);
			}
		} else if (MODE == Z_WRAP) {
			du = __dadd_rn(du, su);
			dv = __dadd_rn(dv, sv);
			dz = __dadd_rn(dz, ((dz > z.wthr) != z.wflip) ? z.wadd : z.shift);
		} else {
			dz = __dadd_rn(dz, z.shift);
		}
		PairS g;
		pair_geometry(du, dv, a0, a1, b0, b1, g);
		unsigned inr = (g.r2 >= rw.lo) & (g.r2 < hi_lane);
		const unsigned zf = (dz >= z.tf), zr = (dz > z.trp);
		unsigned bad = zf ^ zr;  // Pi on an edge: forward and reverse do not land in mirrored slots
		if (MODE == Z_GEN) {
			const unsigned okf = (dz >= z.f_lo) & (dz < z.f_hi), okr = (dz > z.r_lo) & (dz <= z.r_hi);
			bad |= okf ^ okr;
			inr &= okf | okr;
		} else if (MODE == Z_WRAP) {
			bad |= (fabs(dz) > z.zb);
		}
		int slot = (int)zf;
#pragma unroll
		for (int k = 0; k < WRS - 1; k++) slot += (g.r2 >= rw.thr[k]) ? 2 : 0;
		const uint32_t so = (uint32_t)slot * (uint32_t)TP;
		double f0, f1, r0, r1, sw = 0.0;
		lds_v2(f0, f1, acc.af + so * 16u);
		lds_v2(r0, r1, acc.af + so * 16u + AR_OFF);
		const unsigned c0 = lds_u32(acc.ac + so * 4u);
		if (!UNITW) sw = lds_f64(acc.aw + so * 8u);
		const double tf_ = g.crf * g.inv2, tr_ = g.crr * g.inv2;
		double gpf = fma(g.crf, tf_, -1.0);  // cos 2phi = 2 cos^2 - 1
		double gcf = tf_ * fabs(g.srf);      // sin 2phi = 2 cos phi |sin phi|
		const double gpr = fma(g.crr, tr_, -1.0);
		const double gcr = -(tr_ * fabs(g.srr));
		// |sin 2phi| tiny in either ordering -> exact path (integer compare on the high words)
		const int hf = __double2hiint(gcf) & 0x7fffffff, hr = __double2hiint(gcr) & 0x7fffffff;
		bad |= (unsigned)(min(hf, hr) < GC_SMALL_HI);
		// who takes the pair: the galaxy that sees the other one ahead along u (sign and zero test on the bit pattern)
		const int dh = __double2hiint(du);
		const unsigned neg = (unsigned)(dh < 0);
		const unsigned tie = (unsigned)((((unsigned)dh << 1) | (unsigned)__double2loint(du)) == 0u);
		lane_susp |= inr & ((neg & bad) | tie);
		const bool ok = (inr & neg & (bad ^ 1u)) != 0u;
		if (!UNITW) {
			gpf *= cw;
			gcf *= cw;
			sts_f64_if(ok, acc.aw + so * 8u, sw + cw);
		}
		sts_v2_if(ok, acc.af + so * 16u, f0 + gpf, f1 + gcf);
		sts_v2_if(ok, acc.af + so * 16u + AR_OFF, fma(gpr, we, r0), fma(gcr, we, r1));
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

struct FlushCtxS {
	unsigned long long *pcnt;
	double *pddw, *psp, *psc;
	int *flags;
	const double *thr2;  // shared-memory copy of the Pi thresholds
	int n_2, nb, J, num_jk;
};

// Add one (count, sum w, sum e+, sum ex) to row A[jk_shape] and, when the position galaxy lies in another region, to row
// B[jk_pos].  All loads are issued before the first store (one memory latency, not seven).
__device__ __forceinline__ void add_rows(const FlushCtxS &fc, size_t bin, int jk_shape, int jk_pos, unsigned long long cnt, double dw,
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

// Rare path: the pairs pair_loop_sym left out, one lane at a time, with the reference's exact operation sequence for both
// orderings (separation, cos, NaN rule: measure_w_box_jk.py:401-417), added straight into the warp's accumulator copy.
template <bool UNITW, int MODE>
__device__ __noinline__ void slow_pairs_sym(bool lane_susp, uint32_t cb, int n, int periodic, double L, double halfL, double pu,
											double pv, double pl, double a0, double a1, double pe, double pw, int jkS, int jkD,
											const RWindowS rw, double w_hi, const ZSym z, const FlushCtxS fc,
											unsigned long long &nan_pairs, unsigned long long &binned) {
	constexpr uint32_t REC = (uint32_t)sizeof(typename CandRec<UNITW>::type);
	const int lane = threadIdx.x & 31;
	auto sep = [&](double s_, double c_) {  // measure_w_box_jk.py:401-404
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
				PairS g;
				pair_geometry(du, dv, a0, a1, b0, b1, g);
				if (!((g.r2 >= rw.lo) && (g.r2 < w_hi))) continue;
				// this galaxy takes the pair iff the other one is ahead: (d_u, d_v, d_z) lexicographically negative
				if (!(du < 0.0 || (du == 0.0 && (dv < 0.0 || (dv == 0.0 && dz < 0.0))))) continue;
				// the tests pair_loop_sym used to leave the pair out
				const unsigned zf = (dz >= z.tf), zr = (dz > z.trp);
				unsigned bad = zf ^ zr;
				if (MODE == Z_GEN) {
					const unsigned okf = (dz >= z.f_lo) & (dz < z.f_hi), okr = (dz > z.r_lo) & (dz <= z.r_hi);
					bad |= okf ^ okr;
					if (!(okf | okr)) continue;
				} else if (MODE == Z_WRAP) {
					bad |= (fabs(dz) > z.zb);
				}
				const double tf_ = g.crf * g.inv2, tr_ = g.crr * g.inv2;
				const double gcf = tf_ * fabs(g.srf), gcr = -(tr_ * fabs(g.srr));
				const int hf = __double2hiint(gcf) & 0x7fffffff, hr = __double2hiint(gcr) & 0x7fffffff;
				bad |= (unsigned)(min(hf, hr) < GC_SMALL_HI);
				if (!(du == 0.0 || bad)) continue;
				// ---- exact evaluation of both orderings ----
				int rbin = rw.ra;
#pragma unroll
				for (int k = 0; k < WRS - 1; k++) rbin += (g.r2 >= rw.thr[k]) ? 1 : 0;
				const double rp = __dsqrt_rn(g.r2);
				const double ww = pw * cw;
				for (int dir = 0; dir < 2; dir++) {
					const double pi_ = dir ? -dz : dz;  // reverse: sep_ji = -sep_ij exactly
					if (!(pi_ >= fc.thr2[0] && pi_ < fc.thr2[fc.n_2])) continue;
					const int b2 = count_thresholds(pi_, fc.thr2, fc.n_2);
					const double x0 = dir ? -du : du, x1 = dir ? -dv : dv;
					const double c = dir ? __dadd_rn(__dmul_rn(__ddiv_rn(x0, rp), b0), __dmul_rn(__ddiv_rn(x1, rp), b1))
										 : __dadd_rn(__dmul_rn(__ddiv_rn(x0, rp), a0), __dmul_rn(__ddiv_rn(x1, rp), a1));
					double gp = 0.0, gc = 0.0;
					if (fabs(c) <= 1.0) shape_projection(c, gp, gc);
					else nan_pairs++;
					const double amp = dir ? pw * we : pe * cw;  // w_D w_S e_S
					add_rows(fc, (size_t)rbin * fc.n_2 + b2, dir ? jkD : jkS, dir ? jkS : jkD, 1ull, ww, gp * amp, gc * amp);
					binned++;
				}
			}
		}
		__syncwarp();
	}
}

// Flush: fixed-order warp reduction of the private slots into this warp's accumulator copy in HBM.  Lanes are grouped by
// key = (label of the lane, its forward Pi bins, its reverse Pi bins); lane s adds slot s: first all forward sums (rows
// A[label of the lane], B[label of the chunk]), then -- after a warp barrier, because forward and reverse bins of different
// lanes can coincide -- all reverse sums (rows A[label of the chunk], B[label of the lane]).
template <bool UNITW>
__device__ __noinline__ unsigned flush_slots_sym(const FlushCtxS &fc, PrivAccS acc, unsigned long long key, bool dead, double pe,
												 double pw, int ra, int rb, int jkD) {
	static_assert(3 * NSLOT_S <= 32, "one lane per (slot, {forward, reverse, count}) column");
	const int lane = threadIdx.x & 31;
	// Column sums instead of butterflies: lane L < NS sums the forward double2 of slot L over the lanes of the group, lane NS + L
	// the reverse double2, lane 2 NS + L the pair count (and the sum of w_D): 32 serial steps for ALL slots at once, each lane
	// starting at its own lane index (rotated, so that the shared-memory banks are spread) -- a fixed order, about a third of the
	// instructions of five butterfly reductions per slot.  All lanes run the same instruction stream (unused loads are harmless).
	const int role = lane / NSLOT_S, sl = lane - role * NSLOT_S;
	const uint32_t col2 = (acc.af - (uint32_t)lane * 16u) + (uint32_t)sl * TP * 16u + (role == 1 ? AR_OFF : 0u);
	const uint32_t colc = (acc.ac - (uint32_t)lane * 4u) + (uint32_t)sl * TP * 4u;
	const uint32_t colw = (acc.aw - (uint32_t)lane * 8u) + (uint32_t)sl * TP * 8u;
	const uint32_t fac0 = acc.fac - (uint32_t)lane * 16u;
	(void)pe;
	(void)pw;
	__syncwarp();  // the other lanes' private slots are read below: their stores must be visible
	unsigned binned = 0;
	unsigned todo = __ballot_sync(0xffffffffu, !dead);
	while (todo) {
		const int leader = __ffs(todo) - 1;
		const unsigned long long k = __shfl_sync(0xffffffffu, key, leader);
		const unsigned grp = __ballot_sync(0xffffffffu, key == k) & todo;
		double x = 0.0, y = 0.0, dw = 0.0;
		unsigned cnt = 0u;
#pragma unroll 4
		for (int t = 0; t < 32; t++) {
			const int j = (t + lane) & 31;
			const bool in = (grp >> j) & 1u;
			double v0, v1, fe, fw;
			lds_v2(v0, v1, col2 + (uint32_t)j * 16u);
			lds_v2(fe, fw, fac0 + (uint32_t)j * 16u);
			const unsigned c = lds_u32(colc + (uint32_t)j * 4u);
			const double f = in ? (role == 0 ? fe : fw) : 0.0;
			x = fma(v0, f, x);
			y = fma(v1, f, y);
			cnt += in ? c : 0u;
			if (!UNITW) dw = fma(lds_f64(colw + (uint32_t)j * 8u), in ? fw : 0.0, dw);
		}
		// slot L's totals meet in lane L
		const int srcr = (lane + NSLOT_S) & 31, srcc = (lane + 2 * NSLOT_S) & 31;
		const double tf_p = x, tf_c = y;
		const double tr_p = __shfl_sync(0xffffffffu, x, srcr), tr_c = __shfl_sync(0xffffffffu, y, srcr);
		const unsigned tot_cnt = __shfl_sync(0xffffffffu, cnt, srcc);
		const double tot_dw = UNITW ? (double)tot_cnt : __shfl_sync(0xffffffffu, dw, srcc);
		const int kjk = (int)(k >> 32);
		const int kfb0 = (int)((k >> 24) & 0xffu) - 1, kfb1 = (int)((k >> 16) & 0xffu) - 1;
		const int krb0 = (int)((k >> 8) & 0xffu) - 1, krb1 = (int)(k & 0xffu) - 1;
		const int rbin = ra + (lane >> 1);
		const bool mine = lane < NSLOT_S && tot_cnt;
		const int bf = (lane & 1) ? kfb1 : kfb0, br = (lane & 1) ? krb1 : krb0;
		if (mine && (bf < 0 || br < 0 || rbin > rb)) atomicExch(&fc.flags[1], 1);
		const bool go = mine && bf >= 0 && br >= 0 && rbin <= rb;
		if (go) add_rows(fc, (size_t)rbin * fc.n_2 + bf, kjk, jkD, tot_cnt, tot_dw, tf_p, tf_c);
		__syncwarp();
		if (go) {
			add_rows(fc, (size_t)rbin * fc.n_2 + br, jkD, kjk, tot_cnt, tot_dw, tr_p, tr_c);
			binned += 2u * tot_cnt;
		}
		__syncwarp();
		todo &= ~grp;
	}
#pragma unroll
	for (int s_ = 0; s_ < NSLOT_S; s_++) {
		sts_v2(acc.af + (uint32_t)s_ * TP * 16u, 0.0, 0.0);
		sts_v2(acc.af + (uint32_t)s_ * TP * 16u + AR_OFF, 0.0, 0.0);
		if (!UNITW) sts_f64(acc.aw + (uint32_t)s_ * TP * 8u, 0.0);
		sts_u32(acc.ac + (uint32_t)s_ * TP * 4u, 0u);
	}
	return binned;
}

// Everything the chunk consumer needs for one (task, slab, window); lives in the kernel's local memory.
struct S2Ctx {
	// per lane
	double pu, pv, pl, a0, a1, hi_lane, pe, pw;
	ZSym z;
	unsigned long long key;
	int jkS;
	// per (task, slab, window)
	double L, halfL;
	RWindowS rw;
	int rb, periodic, warp_mode;
	PrivAccS acc;
	uint32_t ring_u32;
	const unsigned char *cand;
	unsigned char *ring;
	uint64_t *full;
	FlushCtxS fc;
	// mutable
	uint32_t phase0, phase1;
	int st_issue, cur_label;
	unsigned long long tested, binned, nan_pairs;
};

// Consume one round (see process_round2 in mia_tiled_rppi2.cuh): lane e (bit e of mask) holds a row descriptor = up to two
// contiguous candidate ranges with ONE jackknife label and warp-constant image codes.
template <bool UNITW>
__device__ __forceinline__ void process_round_sym(S2Ctx *cx, int sA, int eA, int sB, int eB, int lab, int codes, unsigned mask) {
	const int lane = threadIdx.x & 31;
	const double L = cx->L, halfL = cx->halfL, pu = cx->pu, pv = cx->pv, pl = cx->pl, a0 = cx->a0, a1 = cx->a1;
	const double hi_lane = cx->hi_lane;
	const RWindowS rw = cx->rw;
	const ZSym z = cx->z;
	const PrivAccS acc = cx->acc;
	const int periodic = cx->periodic;
	typedef typename CandRec<UNITW>::type Rec;
	constexpr int CHS = CandRec<UNITW>::CH;
	const int warp_mode = cx->warp_mode;
	const uint32_t ring_u32 = cx->ring_u32;
	const Rec *cand = reinterpret_cast<const Rec *>(cx->cand);
	Rec *ring = reinterpret_cast<Rec *>(cx->ring);
	uint64_t *full = cx->full;
	uint32_t phase0 = cx->phase0, phase1 = cx->phase1;
	int st_issue = cx->st_issue, cur_label = cx->cur_label;
	unsigned tested = 0, binned = 0;  // per round: < 2^32

	int pend_n = 0, pend_st = 0, pend_label = -1, pend_codes = 0;
	auto consume = [&]() {
		if (pend_label != cur_label) {
			if (cur_label >= 0)
				binned += flush_slots_sym<UNITW>(cx->fc, acc, cx->key, z.dead, cx->pe, cx->pw, rw.ra, cx->rb, cur_label);
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
		if (!z.dead) tested += (unsigned)pend_n;
		const uint32_t cb = ring_u32 + (uint32_t)pend_st * (uint32_t)(CHS * sizeof(Rec));
		const double su = code_shift(cu_, L), sv = code_shift(cv_, L);
		const int mode = (cu_ == 3 || cv_ == 3) ? Z_GEN : ((cu_ | cv_) && warp_mode == Z_PLAIN ? Z_WRAP : warp_mode);
		bool susp;
		if (mode == Z_GEN) susp = pair_loop_sym<UNITW, Z_GEN>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, rw, hi_lane, z, acc);
		else if (mode == Z_WRAP) susp = pair_loop_sym<UNITW, Z_WRAP>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, su, sv, rw, hi_lane, z, acc);
		else susp = pair_loop_sym<UNITW, Z_PLAIN>(cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, 0.0, 0.0, rw, hi_lane, z, acc);
		if (__any_sync(0xffffffffu, susp)) {
			unsigned long long b_ = 0ull;
			if (mode == Z_GEN)
				slow_pairs_sym<UNITW, Z_GEN>(susp, cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, cx->pe, cx->pw, cx->jkS, cur_label,
											 rw, rw.hi, z, cx->fc, cx->nan_pairs, b_);
			else if (mode == Z_WRAP)
				slow_pairs_sym<UNITW, Z_WRAP>(susp, cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, cx->pe, cx->pw, cx->jkS, cur_label,
											  rw, rw.hi, z, cx->fc, cx->nan_pairs, b_);
			else
				slow_pairs_sym<UNITW, Z_PLAIN>(susp, cb, pend_n, periodic, L, halfL, pu, pv, pl, a0, a1, cx->pe, cx->pw, cx->jkS, cur_label,
											   rw, rw.hi, z, cx->fc, cx->nan_pairs, b_);
			cx->binned += b_;
		}
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
				const int rest = en - s, nch = (rest + CHS - 1) / CHS;
				const int n = (rest + nch - 1) / nch;
				if (lane == 0) {
					const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(Rec);
					mbar_expect_tx(&full[st_issue], bytes);
					bulk_load(ring + (size_t)st_issue * CHS, cand + s, bytes, &full[st_issue]);
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

template <bool UNITW>
__device__ __noinline__ void process_round_sym_cold(S2Ctx *cx, int sA, int eA, int sB, int eB, int lab, int codes, unsigned mask) {
	process_round_sym<UNITW>(cx, sA, eA, sB, eB, lab, codes, mask);
}

template <bool UNITW>
__global__ void __launch_bounds__(TP, MIA_S_MIN_CTAS) k_tiled_rppi2s(const TiledArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	const DevParams &P = a.P;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb = P.n_r * P.n_2;
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	const int periodic = P.periodic;
	const double L = P.L, halfL = P.halfL;
	const int nz = a.nz, ncu = P.ncu, ncv = P.ncv, ratio = a.ratio, kk = P.ku;

	// ---- shared memory carve-up ------------------------------------------------------------------------------------
	typedef typename CandRec<UNITW>::type Rec;
	constexpr int CHS = CandRec<UNITW>::CH;
	Rec *ring = reinterpret_cast<Rec *>(smem);  // [warp][stage][CHS]
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + sizeof(Rec) * TW * STAGES * CHS);  // [warp][stage]
	double *thr2_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(full) + 256);
	unsigned char *accbase = reinterpret_cast<unsigned char *>(thr2_s) + 768;
	const uint32_t acc_u32 = smem_u32(accbase);
	Rec *my_ring = ring + (size_t)warp * STAGES * CHS;

	S2Ctx cx;
	cx.acc.af = acc_u32 + (uint32_t)tid * 16u;
	cx.acc.aw = acc_u32 + (uint32_t)NSLOT_S * TP * 32u + (uint32_t)tid * 8u;
	cx.acc.ac = acc_u32 + (uint32_t)NSLOT_S * TP * (UNITW ? 32u : 40u) + (uint32_t)tid * 4u;
	cx.acc.fac = acc_u32 + (uint32_t)NSLOT_S * TP * (UNITW ? 36u : 44u) + (uint32_t)tid * 16u;
	cx.ring_u32 = smem_u32(my_ring);
	cx.ring = reinterpret_cast<unsigned char *>(my_ring);
	cx.full = full + warp * STAGES;
	cx.cand = reinterpret_cast<const unsigned char *>(a.cand);
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
	for (int s = 0; s < NSLOT_S; s++) {
		sts_v2(cx.acc.af + (uint32_t)s * TP * 16u, 0.0, 0.0);
		sts_v2(cx.acc.af + (uint32_t)s * TP * 16u + AR_OFF, 0.0, 0.0);
		if (!UNITW) sts_f64(cx.acc.aw + (uint32_t)s * TP * 8u, 0.0);
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
	cx.fc.flags = a.flags;
	cx.fc.thr2 = thr2_s;
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

	const int n_win = (P.n_r + WRS - 1) / WRS;
	const int n_vr = a.n_lr, vr_cells = ncv / n_vr;
	const int ncv_s = ncv / ratio;
	const int nrows = kk + ratio;  // own rows and the k rows ahead
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
		sts_v2(cx.acc.fac, cx.pe, cx.pw);  // the flush reads the other lanes' factors from shared memory
		__syncwarp();
		cx.jkS = p.jk;
		const int su0 = col / ncv_s;

		const int slab0 = a.task_slab[2 * task], slab1 = a.task_slab[2 * task + 1];
		// per-row v ranges of the two top windows, valid while the same lanes are alive (see the round set-up below)
		float c_vmin0 = 0.f, c_vmax0 = 0.f, c_vmin1 = 0.f, c_vmax1 = 0.f;
		int c_code0 = 0, c_code1 = 0;
		unsigned c_alive0 = 0u, c_alive1 = 0u;  // 0 = nothing cached (a processed slab always has alive != 0)
		for (int s = slab0; s < slab1; s++) {
			const double zlo = a.slab_lo[s], zhi = a.slab_hi[s];
			if (!(zlo <= zhi)) continue;  // empty slab (uniform branch)
			ZSym z = z_window_sym(p.l, zlo, zhi, zp);
			if (!active) z.dead = true;
			if (active && z.err) atomicExch(&a.flags[1], 1);
			const unsigned alive = __ballot_sync(0xffffffffu, !z.dead);
			if (!alive) continue;
			cx.z = z;
			cx.warp_mode = __reduce_max_sync(0xffffffffu, z.dead ? 0 : z.mode);
			cx.key = z.dead ? ~0ull
							: (((unsigned long long)(unsigned)p.jk << 32) | ((unsigned long long)(unsigned)(z.fb0 + 1) << 24) |
							   ((unsigned long long)(unsigned)(z.fb1 + 1) << 16) | ((unsigned long long)(unsigned)(z.rb0 + 1) << 8) |
							   (unsigned long long)(unsigned)(z.rb1 + 1));
			const double bu0 = warp_min_f64(z.dead ? INFINITY : p.u), bu1 = warp_max_f64(z.dead ? -INFINITY : p.u);
			const double bv0 = warp_min_f64(z.dead ? INFINITY : p.v), bv1 = warp_max_f64(z.dead ? -INFINITY : p.v);

			for (int q = 0; q < n_win; q++) {
				// ---- accumulation window q: r bins [ra, rb] counted from the top ------------------------------------------
				const int rb = P.n_r - 1 - q * WRS, ra = (rb - WRS + 1 > 0) ? rb - WRS + 1 : 0;
				RWindowS rw;
				rw.ra = ra;
				rw.lo = P.r2_thr[ra];
				rw.hi = P.r2_thr[rb + 1];
				rw.thr[0] = INFINITY;
#pragma unroll
				for (int t = 0; t < WRS - 1; t++) rw.thr[t] = (ra + 1 + t <= rb) ? P.r2_thr[ra + 1 + t] : INFINITY;
				const double win_hi = rw.hi;
				cx.rw = rw;
				cx.rb = rb;
				cx.hi_lane = z.dead ? -1.0 : rw.hi;  // dead lanes never pass the range test
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
						// ---- ROUND: up to 32 u rows (own rows first, then the rows ahead), one per lane ------------------------
						const int o = rbase + lane;
						int cu = -1;
						if (o < nrows) {
							cu = ratio * su0 + o;
							if (cu >= ncu) cu = periodic ? cu - ncu : -1;
						}
						const long long row = (long long)cu * nz + s;
						// v range of this lane's row within sqrt(r_hi^2 - d_u^2) of some alive shape, from the GEOMETRIC bounds of the
						// row (cell edges, slightly wider than the candidates' bounding box): it depends on the window and on which
						// lanes are alive, not on the slab or the v region, so it is computed once per (task, window) and reused
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
						if (m_simple) process_round_sym<UNITW>(&cx, r_sA, r_eA, r_sB, r_eB, r_lab, r_codes, m_simple);
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
									if (m1) process_round_sym_cold<UNITW>(&cx, c_s, c_e, 0, 0, c_lab, pc, m1);
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
											process_round_sym_cold<UNITW>(&cx, pos, qq, 0, 0, lb, pc, 1u);
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
					cx.binned += flush_slots_sym<UNITW>(cx.fc, cx.acc, cx.key, z.dead, cx.pe, p.w, ra, rb, cx.cur_label);
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

inline int launch_rppi2s(const TiledArgs &a, bool unit_w, int n_ctas, size_t smem, cudaStream_t st) {
	if (unit_w) {
		MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rppi2s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_tiled_rppi2s<true><<<n_ctas, TP, smem, st>>>(a);
	} else {
		MIA_CUDA_CHECK(cudaFuncSetAttribute(k_tiled_rppi2s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_tiled_rppi2s<false><<<n_ctas, TP, smem, st>>>(a);
	}
	return (int)cudaGetLastError();
}

}  // namespace mia

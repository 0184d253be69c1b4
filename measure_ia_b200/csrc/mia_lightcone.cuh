// mia_lightcone.cuh -- brute pair loops of the LIGHT-CONE estimators for sm_100a.
//
// Reference: src/measureia/measure_w_lightcone.py:137-183 (_measure_xi_rp_pi_lightcone_brute), :307-334
// (_count_pairs_xi_rp_pi_lightcone_brute), measure_m_lightcone.py:142-191 / :300-330 (the (r, mu_r) twins).  For every
// POSITION galaxy n the reference forms, against the whole shape sample,
//     Pi  = chi_s - chi_n
//     dx  = ((ra_s - ra_n) / 180 * pi) * chi_n * cos(dec_n / 180 * pi),   dy = ((dec_s - dec_n) / 180 * pi) * chi_n
//     (over_h: dx, dy *= h -- the projected pair only; the 3-D separation of the (r, mu_r) variant keeps the unscaled dx, dy)
//     r_p = sqrt(dx^2 + dy^2),  r = sqrt((dx^2 + dy^2) + Pi^2),  mu_r = Pi / r
//     e_+ = -e cos 2(phi_axis - phi_sep),  e_x = -e sin 2(phi_axis - phi_sep),  phi_sep = arctan2(dy / r_p, dx / r_p)
// bins (r_p, Pi) or (r, mu_r) with the floor-log formula and accumulates w_n w_s {1, e_+, e_x}.  No neighbour search, no
// periodic wrap: O(N_p N_s).
//
// Here: both samples arrive sorted by chi (the host mirror sorts; the library checks).  A CTA owns 128 consecutive position
// galaxies (one per thread) and a segment of the shape sample's chi WINDOW that can reach them (Pi range of the binning,
// or +-r_max): the only cull, exact because every pair inside the window is still tested with the reference's own
// comparisons.  Shape galaxies' sky coordinates are staged through shared memory in tiles for a conservative pre-filter;
// the pairs that pass are compacted per warp (see LC_Q below) before the exact sequence.  Everything that decides a bin uses the
// reference's IEEE operation sequence (__d*_rn, no contraction) against the calibrated thresholds,
// so pair counts are bit-exact.  The shape projection needs no transcendental call per pair:
//     cos 2 phi_sep = (dx^2 - dy^2) / r_p^2,  sin 2 phi_sep = 2 dx dy / r_p^2,
//     e_+ = -(E1 cos 2phi_sep + E2 sin 2phi_sep),  e_x = -(E2 cos 2phi_sep - E1 sin 2phi_sep),
// with E1 = e cos 2phi_axis, E2 = e sin 2phi_axis prepared per shape galaxy by the caller (numpy, the reference's own chain
// theta -> axis -> phi_axis).  Sums go to a per-CTA shared-memory histogram (atomics; binned pairs are a small share of
// the tested ones) and from there to the global result: exact to rounding, not bit-reproducible run to run (like the
// general box kernel); counts are integers and always exact.
//
// Jackknife patches: T[k] = sum over pairs with the position OR the shape galaxy in patch k, so the reference's
// realisation k (both samples without patch k, measure_jackknife.py:116-134) is total - T[k].
#pragma once
#include "mia_common.cuh"

namespace mia {

constexpr int LC_TP = 128;    // threads per CTA = position galaxies per CTA
constexpr int LC_TILE = 128;  // shape galaxies per staged tile

struct LcDev {
	int geom, n_r, n_2, num_patches, shapes, scaled;
	double proj_scale, rp2_cut;
	double r2_thr[MIA_MAX_BINS + 1];
	double thr2[MIA_MAX_BINS + 1];
	double win_lo, win_hi;  // chi_s - chi_n outside [win_lo, win_hi] can never be binned
	double reach;           // > largest separation that can be binned (sqrt of the last r threshold, plus a relative margin)
	double cull_scale;      // factor of the separation components the range test sees: proj_scale for (r_p, Pi), 1 for (r, mu_r)
};

struct LcSampleDev {
	int64_t n;
	const double *ra, *dec, *chi, *cosdec, *w, *e1, *e2;
	const int32_t *patch;
};

struct LcOut {
	unsigned long long *cnt;  // [nb]
	double *ddw, *sp, *sc;    // [nb]
	unsigned long long *jcnt; // [num_patches][nb]
	double *jddw, *jsp;       // [num_patches][nb]
	unsigned long long *stats;
};

__global__ void k_lc_check_sorted(const double *__restrict__ chi, int64_t n, int *__restrict__ flag) {
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i + 1 < n && !(chi[i] <= chi[i + 1])) atomicExch(flag, 1);  // also catches NaN
	if (i < n && !(chi[i] == chi[i])) atomicExch(flag, 1);
}

// One pair through the reference's exact operation sequence; accumulates into the CTA's shared histogram (and the global
// per-patch rows).  Returns true when the pair was binned.
struct LcAcc {
	double *s_ddw, *s_sp, *s_sc;
	unsigned int *s_cnt;
	int nb;
};

template <int GEOM, bool SHAPES>
__device__ __forceinline__ bool lc_pair(const LcDev &P, const LcOut &O, const LcAcc &A, double ra_n, double dec_n, double chi_n,
										double cd_n, double w_n, int patch_n, double ra_s, double dec_s, double chi_s, double w_s,
										double e1_s, double e2_s, int patch_s) {
	const double los = __dsub_rn(chi_s, chi_n);  // measure_w_lightcone.py:139
	if (GEOM == MIA_GEOM_RPPI) {
		if (!(los >= P.thr2[0] && los < P.thr2[P.n_2])) return false;  // :160-161
	}
	const double dra = __dmul_rn(__ddiv_rn(__dsub_rn(ra_s, ra_n), 180.0), 3.141592653589793);    // :140
	const double ddec = __dmul_rn(__ddiv_rn(__dsub_rn(dec_s, dec_n), 180.0), 3.141592653589793);  // :141
	const double dx = __dmul_rn(__dmul_rn(dra, chi_n), cd_n);                                      // :142
	const double dy = __dmul_rn(ddec, chi_n);                                                        // :143
	const double px = P.scaled ? __dmul_rn(dx, P.proj_scale) : dx, py = P.scaled ? __dmul_rn(dy, P.proj_scale) : dy;  // :145-146
	const double rp2 = __dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py));  // :147
	double s;
	int bin2;
	if (GEOM == MIA_GEOM_RPPI) {
		s = rp2;
		if (!(s >= P.r2_thr[0] && s < P.r2_thr[P.n_r])) return false;
		bin2 = count_thresholds(los, P.thr2, P.n_2);
	} else {
		if (!(rp2 > P.rp2_cut)) return false;  // measure_m_lightcone.py:172
		// the 3-D separation keeps the UNSCALED dx, dy (:150 builds it before `projected_sep *= h`)
		s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(los, los));  // :154
		if (!(s >= P.r2_thr[0] && s < P.r2_thr[P.n_r])) return false;
		const double mu = __ddiv_rn(los, __dsqrt_rn(s));  // :157
		bin2 = count_thresholds(mu, P.thr2, P.n_2);
	}
	const int b = count_thresholds(s, P.r2_thr, P.n_r) * P.n_2 + bin2;
	const double ww = w_n * w_s;
	double tp = 0.0, tc = 0.0;
	if (SHAPES) {
		// rp2 == 0 cannot be binned ((r_p, Pi): r_p >= r_min > 0; (r, mu_r): r_p^2 > rp2_cut >= 0), so the reference's
		// NaN -> 0 rule (:153-154) never fires on a binned pair
		const double inv = 1.0 / rp2;
		const double c2 = (px * px - py * py) * inv, s2 = 2.0 * px * py * inv;
		tp = -ww * (e1_s * c2 + e2_s * s2);
		tc = -ww * (e2_s * c2 - e1_s * s2);
	}
	atomicAdd(&A.s_cnt[b], 1u);
	atomicAdd(&A.s_ddw[b], ww);
	if (SHAPES) {
		atomicAdd(&A.s_sp[b], tp);
		atomicAdd(&A.s_sc[b], tc);
	}
	if (P.num_patches > 0) {
		size_t row = (size_t)patch_s * A.nb + b;
		atomicAdd(&O.jcnt[row], 1ull);
		atomicAdd(&O.jddw[row], ww);
		if (SHAPES) atomicAdd(&O.jsp[row], tp);
		if (patch_n != patch_s) {
			row = (size_t)patch_n * A.nb + b;
			atomicAdd(&O.jcnt[row], 1ull);
			atomicAdd(&O.jddw[row], ww);
			if (SHAPES) atomicAdd(&O.jsp[row], tp);
		}
	}
	return true;
}

// Shared memory: shape tile (6 doubles + patch) | the CTA's position galaxies (5 doubles + patch) | histogram | per-warp queues.
// Pairs that pass the pre-filter are rare (about one lane in a hundred per iteration) and the exact sequence is ~150
// instructions, so running it in place would drag the whole warp through it for one or two lanes.  Instead the passing
// (lane, shape index) pairs go to a per-warp ring in shared memory, and whenever 32 are waiting every lane takes one:
// the exact sequence always runs with full warps (the shape galaxy is re-read from global memory, an L2 hit).
constexpr int LC_Q = 64;  // ring entries per warp (at most 31 waiting + 32 new)

template <int GEOM, bool SHAPES>
__global__ void __launch_bounds__(LC_TP) k_lightcone(const LcDev P, const LcSampleDev D, const LcSampleDev S, int64_t p_begin,
													 int64_t p_end, int n_seg, LcOut O) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int nb = P.n_r * P.n_2;
	double2 *t_sky = reinterpret_cast<double2 *>(smem_raw);  // {dec, ra} of the staged shape galaxies: one 16-byte load per pair
	double *p_ra = reinterpret_cast<double *>(t_sky + LC_TILE), *p_dec = p_ra + LC_TP, *p_chi = p_dec + LC_TP, *p_cd = p_chi + LC_TP, *p_w = p_cd + LC_TP;
	double *s_ddw = p_w + LC_TP, *s_sp = s_ddw + nb, *s_sc = s_sp + nb;
	unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_sc + nb);
	int *p_patch = reinterpret_cast<int *>(s_cnt + nb);
	int *q_j = p_patch + LC_TP;                                               // [warps][LC_Q] shape index
	unsigned char *q_lane = reinterpret_cast<unsigned char *>(q_j + (LC_TP / 32) * LC_Q);  // [warps][LC_Q] lane of the position galaxy
	__shared__ long long win[2];

	for (int b = threadIdx.x; b < nb; b += blockDim.x) {
		s_ddw[b] = 0.0;
		s_sp[b] = 0.0;
		s_sc[b] = 0.0;
		s_cnt[b] = 0u;
	}
	const int64_t blk0 = p_begin + (int64_t)blockIdx.x * LC_TP;
	const int64_t blk1 = (blk0 + LC_TP < p_end) ? blk0 + LC_TP : p_end;
	if (threadIdx.x == 0) {
		// shape galaxies that can reach this block: chi_s in [chi(first) + win_lo, chi(last) + win_hi] (positions ascend in chi),
		// widened by a relative slack so that the rounding of the subtraction can never exclude a pair the exact test accepts
		const double c0 = D.chi[blk0], c1 = D.chi[blk1 - 1];
		const double lo = c0 + P.win_lo - 1e-9 * (fabs(c0) + fabs(P.win_lo)) - 1e-300;
		const double hi = c1 + P.win_hi + 1e-9 * (fabs(c1) + fabs(P.win_hi)) + 1e-300;
		long long a = 0, b = S.n;
		if (lo == lo && lo > -INFINITY) {
			while (a < b) {  // first index with chi_s >= lo
				const long long m = (a + b) >> 1;
				if (S.chi[m] >= lo) b = m;
				else a = m + 1;
			}
		}
		win[0] = a;
		a = win[0];
		b = S.n;
		if (hi == hi && hi < INFINITY) {
			while (a < b) {  // first index with chi_s > hi
				const long long m = (a + b) >> 1;
				if (S.chi[m] > hi) b = m;
				else a = m + 1;
			}
			win[1] = a;
		} else {
			win[1] = S.n;
		}
	}
	const int64_t n = blk0 + threadIdx.x;
	const bool active = n < blk1;
	double ra_n = 0.0, dec_n = 0.0, chi_n = 0.0, cd_n = 0.0, w_n = 0.0;
	int patch_n = 0;
	if (active) {
		ra_n = D.ra[n];
		dec_n = D.dec[n];
		chi_n = D.chi[n];
		cd_n = D.cosdec[n];
		w_n = D.w ? D.w[n] : 1.0;
		patch_n = D.patch ? D.patch[n] : 0;
	}
	p_ra[threadIdx.x] = ra_n;
	p_dec[threadIdx.x] = dec_n;
	p_chi[threadIdx.x] = chi_n;
	p_cd[threadIdx.x] = cd_n;
	p_w[threadIdx.x] = w_n;
	p_patch[threadIdx.x] = patch_n;
	__syncthreads();
	const long long w0 = win[0], w1 = win[1];

	// conservative pre-filter on the two sky offsets (a few ulp of error against a 1e-9 margin in `reach`): a pair whose dec
	// or ra offset ALONE exceeds the largest binnable separation is skipped before the exact operation sequence is paid for
	const double k_dec = (3.141592653589793 / 180.0) * chi_n * P.cull_scale;
	const double lim_dec = (k_dec > 0.0) ? P.reach / k_dec : INFINITY;  // degrees
	const double lim_ra = (k_dec * fabs(cd_n) > 0.0) ? P.reach / (k_dec * fabs(cd_n)) : INFINITY;
	const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
	int *my_qj = q_j + (threadIdx.x >> 5) * LC_Q;
	unsigned char *my_ql = q_lane + (threadIdx.x >> 5) * LC_Q;
	int q_head = 0, q_n = 0;  // warp-uniform
	LcAcc A = {s_ddw, s_sp, s_sc, s_cnt, nb};
	unsigned long long tested = 0, binned = 0;

	auto take_entry = [&](int slot) {  // this lane evaluates the queued pair in ring slot `slot`
		const int src = wbase + (int)my_ql[slot];
		const long long j = (long long)my_qj[slot];
		const bool ok = lc_pair<GEOM, SHAPES>(P, O, A, p_ra[src], p_dec[src], p_chi[src], p_cd[src], p_w[src], p_patch[src], S.ra[j],
											  S.dec[j], S.chi[j], S.w ? S.w[j] : 1.0, SHAPES ? S.e1[j] : 0.0, SHAPES ? S.e2[j] : 0.0,
											  S.patch ? S.patch[j] : 0);
		binned += ok ? 1ull : 0ull;
	};

	const long long n_tiles = (w1 - w0 + LC_TILE - 1) / LC_TILE;
	for (long long t = blockIdx.y; t < n_tiles; t += n_seg) {
		const long long j0 = w0 + t * LC_TILE;
		const int cnt = (int)((w1 - j0 < LC_TILE) ? (w1 - j0) : LC_TILE);
		__syncthreads();  // the previous tile has been consumed
		if ((int)threadIdx.x < cnt) t_sky[threadIdx.x] = make_double2(S.dec[j0 + threadIdx.x], S.ra[j0 + threadIdx.x]);
		__syncthreads();
#pragma unroll 1
		for (int k = 0; k < cnt; k++) {
			const double2 sk = t_sky[k];
			// branch-free (a NaN coordinate fails both comparisons: the reference's range mask rejects such a pair too)
			const bool pass = active & (fabs(sk.x - dec_n) <= lim_dec) & (fabs(sk.y - ra_n) <= lim_ra);
			const unsigned m = __ballot_sync(0xffffffffu, pass);
			if (!m) continue;
			if (pass) {
				const int slot = (q_head + q_n + __popc(m & ((1u << lane) - 1u))) & (LC_Q - 1);
				my_qj[slot] = (int)(j0 + k);
				my_ql[slot] = (unsigned char)lane;
				tested++;
			}
			q_n += __popc(m);
			__syncwarp();
			if (q_n >= 32) {
				take_entry((q_head + lane) & (LC_Q - 1));
				q_head = (q_head + 32) & (LC_Q - 1);
				q_n -= 32;
				__syncwarp();
			}
		}
	}
	if (lane < q_n) take_entry((q_head + lane) & (LC_Q - 1));  // what is left in the ring
	__syncthreads();
	for (int b = threadIdx.x; b < nb; b += blockDim.x) {
		if (s_cnt[b]) {
			atomicAdd(&O.cnt[b], (unsigned long long)s_cnt[b]);
			atomicAdd(&O.ddw[b], s_ddw[b]);
			if (SHAPES) {
				atomicAdd(&O.sp[b], s_sp[b]);
				atomicAdd(&O.sc[b], s_sc[b]);
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		tested += __shfl_down_sync(0xffffffffu, tested, o);
		binned += __shfl_down_sync(0xffffffffu, binned, o);
	}
	if (lane == 0) {
		atomicAdd(&O.stats[0], tested);
		atomicAdd(&O.stats[1], binned);
	}
}

inline size_t lightcone_smem_bytes(int nb) {
	return sizeof(double) * (2 * LC_TILE + 5 * LC_TP) + (size_t)nb * (3 * sizeof(double) + sizeof(unsigned int)) + sizeof(int) * LC_TP +
		   (size_t)(LC_TP / 32) * LC_Q * (sizeof(int) + 1);
}

}  // namespace mia

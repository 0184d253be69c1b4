// mia_general.cuh -- the GENERAL pair kernel: one thread per shape galaxy, reference-exact evaluation of every
// candidate in the neighbouring cells, atomics for accumulation.
//
// Role: (1) handles every configuration (tiny boxes where the cell grid degenerates to one cell, many bins,
// non-periodic boxes, ...); (2) is the independent on-GPU cross-check of the tiled kernel.  It is not the fast path:
// fp64 sums are accumulated with atomics, so they are exact to rounding but not bit-reproducible run to run (pair
// counts are integers and always exact).
#pragma once
#include "mia_common.cuh"

namespace mia {

struct GeneralSmem {
	unsigned int *cnt;
	double *ddw, *sp, *sc;
};

template <int GEOM>
__global__ void __launch_bounds__(128) k_general(DevParams P, Grid G, const Prim *__restrict__ prim, int64_t s_begin,
												 int64_t s_end, Accum A) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int nb = P.n_r * P.n_2;
	const bool use_smem = (P.num_jk == 0);
	double *s_ddw = reinterpret_cast<double *>(smem_raw);
	double *s_sp = s_ddw + nb, *s_sc = s_sp + nb;
	unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_sc + nb);
	if (use_smem) {
		for (int b = threadIdx.x; b < nb; b += blockDim.x) {
			s_ddw[b] = 0.0;
			s_sp[b] = 0.0;
			s_sc[b] = 0.0;
			s_cnt[b] = 0u;
		}
		__syncthreads();
	}
	const int J = P.num_jk > 0 ? P.num_jk : 1;
	unsigned long long tested = 0, binned = 0, nanpairs = 0;

	int64_t i = s_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i < s_end) {
		const Prim p = prim[i];
		const int cu = cell_index(p.u, P.inv_cu, P.ncu), cv = cell_index(p.v, P.inv_cv, P.ncv),
				  cl = cell_index(p.l, P.inv_cl, P.ncl);
		// neighbour ranges: either relative offsets -k..k (with periodic wrap / clipping) or the whole axis
		const bool all_u = 2 * P.ku + 1 >= P.ncu, all_v = 2 * P.kv + 1 >= P.ncv, all_l = 2 * P.kl + 1 >= P.ncl;
		const int u0 = all_u ? 0 : cu - P.ku, u1 = all_u ? P.ncu - 1 : cu + P.ku;
		const int v0 = all_v ? 0 : cv - P.kv, v1 = all_v ? P.ncv - 1 : cv + P.kv;
		const int l0 = all_l ? 0 : cl - P.kl, l1 = all_l ? P.ncl - 1 : cl + P.kl;
		const double pw_e = p.w * p.e;
		for (int iu = u0; iu <= u1; iu++) {
			int nu = iu;
			if (nu < 0 || nu >= P.ncu) {
				if (!P.periodic) continue;
				nu = (nu + P.ncu) % P.ncu;
			}
			for (int iv = v0; iv <= v1; iv++) {
				int nv = iv;
				if (nv < 0 || nv >= P.ncv) {
					if (!P.periodic) continue;
					nv = (nv + P.ncv) % P.ncv;
				}
				for (int il = l0; il <= l1; il++) {
					int nl = il;
					if (nl < 0 || nl >= P.ncl) {
						if (!P.periodic) continue;
						nl = (nl + P.ncl) % P.ncl;
					}
					const int64_t cell = ((int64_t)nu * P.ncv + nv) * P.ncl + nl;
					const int64_t j0 = G.cell_start[cell], j1 = G.cell_start[cell + 1];
					for (int64_t j = j0; j < j1; j++) {
						const Cand c = G.cand[j];
						PairResult r;
						tested++;
						if (!eval_pair_exact<GEOM>(P, p.u, p.v, p.l, p.a0, p.a1, c.u, c.v, c.l, r)) continue;
						binned++;
						nanpairs += r.nan_rule ? 1 : 0;
						const int b = r.rbin * P.n_2 + r.bin2;
						const double ww = c.w * p.w;
						const double tp = c.w * pw_e * r.gp, tc = c.w * pw_e * r.gc;
						if (use_smem) {
							atomicAdd(&s_cnt[b], 1u);
							atomicAdd(&s_ddw[b], ww);
							atomicAdd(&s_sp[b], tp);
							atomicAdd(&s_sc[b], tc);
							if (A.var) atomicAdd(&A.var[b], tp * tp);  // brute variants' `variance`, measure_w_box_jk.py:196
						} else {
							const size_t ra = (size_t)p.jk * nb + b;
							atomicAdd(&A.cnt[ra], 1ull);
							atomicAdd(&A.ddw[ra], ww);
							atomicAdd(&A.sp[ra], tp);
							atomicAdd(&A.sc[ra], tc);
							if (A.var) atomicAdd(&A.var[b], tp * tp);
							const int jd = G.cand_jk[j];
							if (jd != p.jk) {
								const size_t rb = (size_t)(J + jd) * nb + b;
								atomicAdd(&A.cnt[rb], 1ull);
								atomicAdd(&A.ddw[rb], ww);
								atomicAdd(&A.sp[rb], tp);
								atomicAdd(&A.sc[rb], tc);
							}
						}
					}
				}
			}
		}
	}
	if (use_smem) {
		__syncthreads();
		for (int b = threadIdx.x; b < nb; b += blockDim.x) {
			if (s_cnt[b]) {
				atomicAdd(&A.cnt[b], (unsigned long long)s_cnt[b]);
				atomicAdd(&A.ddw[b], s_ddw[b]);
				atomicAdd(&A.sp[b], s_sp[b]);
				atomicAdd(&A.sc[b], s_sc[b]);
			}
		}
	}
	// statistics: warp-reduce, one atomic per warp
	for (int o = 16; o > 0; o >>= 1) {
		tested += __shfl_down_sync(0xffffffffu, tested, o);
		binned += __shfl_down_sync(0xffffffffu, binned, o);
		nanpairs += __shfl_down_sync(0xffffffffu, nanpairs, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&A.stats[0], tested);
		atomicAdd(&A.stats[1], binned);
		atomicAdd(&A.stats[2], nanpairs);
	}
}

}  // namespace mia

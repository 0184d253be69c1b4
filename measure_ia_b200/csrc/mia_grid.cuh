// mia_grid.cuh -- GPU cell-list build: cell-key radix sort + cell offsets (replaces the reference's
// scipy.spatial.KDTree(positions[:, not_LOS], boxsize=L), measure_w_box_jk.py:386,394-397 / measure_m_box_jk.py:403).
//
// Layout in HBM after the build (all in the caller's workspace):
//   Cand  cand[nD]        position sample, canonical axes, sorted by key = cell * jk_rows + jk_label   (32 B each)
//   int32 cand_jk[nD]     jackknife label per sorted candidate
//   Prim  prim[nS]        shape sample, same sort                                                       (64 B each)
//   int64 cell_start[ncell+1]  (candidates)   int64 prim_cell_start[ncell+1]  (primaries)
// The sort is a stable LSD radix sort (CUB), so the order inside a cell is the caller's index order: the
// accumulation order, and therefore every fp64 sum of the tiled kernel, is reproducible run to run.
#pragma once
#include <cub/cub.cuh>
#include "mia_common.cuh"

namespace mia {

struct GridDims {
	int ncu, ncv, ncl;
	double inv_cu, inv_cv, inv_cl;
	int jk_rows;  // max(num_jk, 1)
	int sub;      // shape sample only: each cell's galaxies are ordered by sub x sub projected sub-cell (compact warps)
	int order;    // 0: cell = (cu, cv, cl) -- the slabs of a column are contiguous; 1: cell = (cu, cl, cv) -- the cells of a
	              // (u row, slab) are contiguous along v (row-streaming (r_p, Pi) kernel)
	int64_t ncell() const { return (int64_t)ncu * ncv * ncl; }
	int64_t nsorted_cells() const { return ncell() * (sub > 1 ? sub * sub : 1); }
};

__global__ void k_make_keys(const double *__restrict__ pos, const int32_t *__restrict__ jk, int64_t n, int nl0, int nl1,
							int los, GridDims g, double L, uint32_t *__restrict__ keys, int32_t *__restrict__ idx,
							int *__restrict__ range_err) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	double u = pos[3 * i + nl0], v = pos[3 * i + nl1], l = pos[3 * i + los];
	// the reference's periodic KDTree requires 0 <= x < L (scipy raises otherwise); NaNs fail here as well
	if (!(u >= 0.0 && u < L && v >= 0.0 && v < L && l >= 0.0 && l < L)) atomicExch(range_err, 1);
	int cu = cell_index(u, g.inv_cu, g.ncu), cv = cell_index(v, g.inv_cv, g.ncv), cl = cell_index(l, g.inv_cl, g.ncl);
	uint32_t cell = g.order ? (uint32_t)((cu * g.ncl + cl) * g.ncv + cv) : (uint32_t)((cu * g.ncv + cv) * g.ncl + cl);
	if (g.sub > 1) {
		int su = (int)((u * g.inv_cu - cu) * g.sub), sv = (int)((v * g.inv_cv - cv) * g.sub);
		su = su < 0 ? 0 : (su >= g.sub ? g.sub - 1 : su);
		sv = sv < 0 ? 0 : (sv >= g.sub ? g.sub - 1 : sv);
		cell = cell * (uint32_t)(g.sub * g.sub) + (uint32_t)(su * g.sub + sv);
	}
	uint32_t lab = jk ? (uint32_t)jk[i] : 0u;
	keys[i] = cell * (uint32_t)g.jk_rows + lab;
	idx[i] = (int32_t)i;
}

__global__ void k_gather_cand(const double *__restrict__ pos, const double *__restrict__ w,
							  const int32_t *__restrict__ jk, const int32_t *__restrict__ idx, int64_t n, int nl0,
							  int nl1, int los, Cand *__restrict__ out, int32_t *__restrict__ out_jk) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	int64_t s = idx[i];
	Cand c;
	c.u = pos[3 * s + nl0];
	c.v = pos[3 * s + nl1];
	c.l = pos[3 * s + los];
	c.w = w ? w[s] : 1.0;
	out[i] = c;
	out_jk[i] = jk ? jk[s] : 0;
}

__global__ void k_gather_prim(const double *__restrict__ pos, const double *__restrict__ w,
							  const int32_t *__restrict__ jk, const double *__restrict__ axis,
							  const double *__restrict__ e, const int32_t *__restrict__ idx, int64_t n, int nl0, int nl1,
							  int los, Prim *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	int64_t s = idx[i];
	Prim p;
	p.u = pos[3 * s + nl0];
	p.v = pos[3 * s + nl1];
	p.l = pos[3 * s + los];
	p.w = w ? w[s] : 1.0;
	p.a0 = axis[2 * s];
	p.a1 = axis[2 * s + 1];
	p.e = e[s];
	p.jk = jk ? jk[s] : 0;
	p.orig = (int32_t)s;
	out[i] = p;
}

// cell_start[c] = first sorted index whose cell >= c; cell_start[ncell] = n.
__global__ void k_cell_start(const uint32_t *__restrict__ keys, int64_t n, int jk_rows, int64_t ncell,
							 int64_t *__restrict__ cell_start) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i > n) return;
	int64_t c_prev = (i == 0) ? -1 : (int64_t)(keys[i - 1] / (uint32_t)jk_rows);
	int64_t c = (i == n) ? ncell : (int64_t)(keys[i] / (uint32_t)jk_rows);
	for (int64_t cc = c_prev + 1; cc <= c; cc++) cell_start[cc] = i;
}

struct SortScratch {
	uint32_t *keys_in, *keys_out;
	int32_t *idx_in, *idx_out;
	void *cub_tmp;
	size_t cub_bytes;
};

inline size_t cub_sort_bytes(int64_t n) {
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
									(const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
	return bytes;
}

// Sort one sample by cell key.  On return sc.keys_out / sc.idx_out hold the sorted keys and source indices.
inline int sort_by_cell(const double *pos, const int32_t *jk, int64_t n, int nl0, int nl1, int los, const GridDims &g,
						double L, SortScratch &sc, int key_bits, int *range_err, cudaStream_t st) {
	if (n == 0) return 0;
	const int T = 256;
	const unsigned B = (unsigned)((n + T - 1) / T);
	k_make_keys<<<B, T, 0, st>>>(pos, jk, n, nl0, nl1, los, g, L, sc.keys_in, sc.idx_in, range_err);
	MIA_CUDA_CHECK(cudaGetLastError());
	MIA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sc.cub_tmp, sc.cub_bytes, sc.keys_in, sc.keys_out, sc.idx_in,
												   sc.idx_out, (int)n, 0, key_bits, st));
	return 0;
}

}  // namespace mia

// mia_api.cu -- C ABI of libmia_b200.so (see include/mia_b200.h) and the host-side orchestration:
// workspace carving, cell-list build, kernel selection, fixed-order final reduction.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "mia_common.cuh"
#include "mia_general.cuh"
#include "mia_grid.cuh"
#include "mia_lightcone.cuh"
#include "mia_tiled.cuh"
#include "mia_tiled_rmu.cuh"
#include "mia_tiled_rppi2.cuh"
#include "mia_tiled_rppi2s.cuh"

using namespace mia;

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// CUDA events destroyed on every exit path
struct CudaEvents {
	cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
	int create(int n) {
		for (int i = 0; i < n && i < 4; i++) MIA_CUDA_CHECK(cudaEventCreate(&e[i]));
		return 0;
	}
	~CudaEvents() {
		for (int i = 0; i < 4; i++)
			if (e[i]) cudaEventDestroy(e[i]);
	}
};

constexpr int RED_GROUPS = 32;  // groups of accumulator copies summed in parallel (stage 1 of the final reduction)

struct Plan {
	// grid
	GridDims g;
	int ku, kv, kl;
	int kernel;       // resolved MIA_KERNEL_*
	int key_bits;
	int n_partials;   // accumulator copies (1 for the general kernel, one per CTA for the tiled kernel)
	int cand_bytes;   // bytes per sorted candidate record (Cand, or CandS when the symmetric kernel may run)
	int rows;         // 2 * max(num_jk, 1)
	int nb;
	TiledConfig tiled;
	// workspace offsets
	size_t off_cand, off_cand_jk, off_prim, off_cell_start, off_prim_cell_start, off_keys_in, off_keys_out, off_idx_in,
		off_idx_out, off_cub, cub_bytes, off_red_cnt, off_red_f, off_cnt, off_ddw, off_sp, off_sc, off_var, off_stats, off_flags,
		off_tiled, total;
};

int ilog2_ceil(uint64_t x) {
	int b = 0;
	while ((1ull << b) < x) b++;
	return b;
}

int validate(const mia_params *p) {
	if (!p || p->abi_version != MIA_ABI_VERSION) return MIA_ERR_ARG;
	if (p->n_r < 1 || p->n_r > MIA_MAX_BINS || p->n_2 < 1 || p->n_2 > MIA_MAX_BINS) return MIA_ERR_ARG;
	if (p->los < 0 || p->los > 2 || p->num_jk < 0 || !(p->boxsize > 0.0) || !(p->r_search > 0.0)) return MIA_ERR_ARG;
	if (p->geometry != MIA_GEOM_RPPI && p->geometry != MIA_GEOM_RMU) return MIA_ERR_ARG;
	if (!p->r2_thr_host || !p->thr2_host) return MIA_ERR_ARG;
	return MIA_OK;
}

// Cell grid for the general kernel: cells at least one search radius wide, 3 x 3 (x 3) neighbourhood.
void plan_general_grid(const mia_params *p, Plan &pl) {
	const double reach = p->r_search * (1.0 + 1e-6);
	int nc = (int)floor(p->boxsize / reach);
	const int cap = (p->geometry == MIA_GEOM_RPPI) ? 1024 : 160;
	if (nc > cap) nc = cap;
	if (nc < 3) nc = 1;
	pl.g.ncu = pl.g.ncv = nc;
	pl.g.ncl = (p->geometry == MIA_GEOM_RMU) ? nc : 1;
	pl.g.inv_cu = pl.g.inv_cv = nc / p->boxsize;
	pl.g.inv_cl = pl.g.ncl / p->boxsize;
	pl.ku = pl.kv = 1;
	pl.kl = (p->geometry == MIA_GEOM_RMU) ? 1 : 0;
}

int make_plan(const mia_params *p, int64_t nD, int64_t nS, Plan &pl) {
	memset(&pl, 0, sizeof(pl));
	pl.nb = p->n_r * p->n_2;
	const int J = p->num_jk > 0 ? p->num_jk : 1;
	pl.rows = 2 * J;
	pl.g.jk_rows = J;
	pl.kernel = p->kernel;
	const bool ordered_only = (pl.kernel == MIA_KERNEL_TILED_ORDERED);
	if (ordered_only) pl.kernel = MIA_KERNEL_TILED;
	if (pl.kernel == MIA_KERNEL_AUTO || pl.kernel == MIA_KERNEL_TILED) {
		// the tiled grids are fine (cells ~ r_max / 4): their sort keys must fit 31 bits
		if (plan_tiled(p, nD, nS, pl.g, pl.ku, pl.kv, pl.kl, pl.tiled) &&
			(uint64_t)pl.g.ncell() * 4ull * (uint64_t)J <= (1ull << 31)) {
			pl.kernel = MIA_KERNEL_TILED;
		} else {
			if (pl.kernel == MIA_KERNEL_TILED) return MIA_ERR_UNSUPPORTED;
			pl.kernel = MIA_KERNEL_GENERAL;
			// not silent: the general kernel is ~30x slower (one thread per shape galaxy, atomics)
			static bool warned = false;
			if (!warned && nD * nS > 0) {
				warned = true;
				fprintf(stderr,
						"libmia_b200: warning: this configuration (%s, %d x %d bins, %d jackknife regions) is outside what the "
						"tiled kernels cover; falling back to the general kernel (about 30x slower)\n",
						p->geometry == MIA_GEOM_RPPI ? "(r_p, Pi)" : "(r, mu_r)", p->n_r, p->n_2, p->num_jk);
			}
		}
	}
	if (pl.kernel == MIA_KERNEL_GENERAL) {
		plan_general_grid(p, pl);
		pl.n_partials = 1;
		pl.tiled.v2 = 0;
		pl.tiled.ratio = 1;
	} else {
		pl.n_partials = pl.tiled.n_partials;
		pl.g.order = pl.tiled.v2 ? 1 : 0;  // row-streaming (r_p, Pi) kernel: candidates sorted by (u row, slab, v cell)
		if (ordered_only || p->variance || env_int("MIA_SYM", 1) == 0) pl.tiled.sym_ok = 0;
	}
	// the symmetric auto-correlation kernel needs 64-byte candidate records; whether the two samples really are the same
	// catalogue is only known at call time, so the workspace is sized for it whenever the sizes agree
	pl.cand_bytes = (pl.kernel == MIA_KERNEL_TILED && pl.tiled.sym_ok && nD == nS) ? (int)sizeof(CandSW) : (int)sizeof(Cand);
	const uint64_t nkeys = (uint64_t)pl.g.ncell() * 4ull * (uint64_t)J;  // x4: the shape sample's sub-cell ordering
	if (nkeys > (1ull << 31)) return MIA_ERR_UNSUPPORTED;
	pl.key_bits = ilog2_ceil(nkeys > 1 ? nkeys : 2);
	const int64_t nmax = nD > nS ? nD : nS;
	if (nmax >= (1ll << 31)) return MIA_ERR_UNSUPPORTED;

	size_t o = 0;
	auto take = [&](size_t bytes) {
		size_t at = o;
		o = align_up(o + bytes);
		return at;
	};
	pl.off_cand = take((size_t)pl.cand_bytes * (size_t)nD);
	pl.off_cand_jk = take(sizeof(int32_t) * (size_t)nD);
	pl.off_prim = take(sizeof(Prim) * (size_t)nS);
	pl.off_cell_start = take(sizeof(int64_t) * (size_t)(pl.g.ncell() + 1));
	pl.off_prim_cell_start = take(sizeof(int64_t) * (size_t)(pl.g.ncell() * 4 + 1));
	pl.off_keys_in = take(sizeof(uint32_t) * (size_t)nmax);
	pl.off_keys_out = take(sizeof(uint32_t) * (size_t)nmax);
	pl.off_idx_in = take(sizeof(int32_t) * (size_t)nmax);
	pl.off_idx_out = take(sizeof(int32_t) * (size_t)nmax);
	pl.cub_bytes = cub_sort_bytes(nmax > 0 ? nmax : 1);
	pl.off_cub = take(pl.cub_bytes);
	const size_t acc = (size_t)pl.n_partials * pl.rows * pl.nb;
	pl.off_red_cnt = take(pl.n_partials > 1 ? sizeof(unsigned long long) * RED_GROUPS * (size_t)pl.rows * pl.nb : 0);
	pl.off_red_f = take(pl.n_partials > 1 ? sizeof(double) * 3 * RED_GROUPS * (size_t)pl.rows * pl.nb : 0);
	pl.off_cnt = take(sizeof(unsigned long long) * acc);
	pl.off_ddw = take(sizeof(double) * acc);
	pl.off_sp = take(sizeof(double) * acc);
	pl.off_sc = take(sizeof(double) * acc);
	pl.off_var = take(p->variance ? sizeof(double) * (size_t)pl.n_partials * pl.nb : 0);
	pl.off_stats = take(sizeof(unsigned long long) * 8);
	pl.off_flags = take(sizeof(int) * 8);
	pl.off_tiled = take(pl.kernel == MIA_KERNEL_TILED ? tiled_workspace_bytes(pl.tiled, pl.g, nD, nS) : 0);
	pl.total = o;
	return MIA_OK;
}

void fill_dev_params(const mia_params *p, const Plan &pl, DevParams &P) {
	memset(&P, 0, sizeof(P));
	P.geom = p->geometry;
	P.n_r = p->n_r;
	P.n_2 = p->n_2;
	P.los = p->los;
	P.periodic = p->periodic ? 1 : 0;
	P.num_jk = p->num_jk;
	P.L = p->boxsize;
	P.halfL = p->boxsize / 2.0;  // L_0p5 = boxsize / 2. (Sim_info.py:79)
	P.rp2_cut = p->rp2_cut;
	for (int b = 0; b <= p->n_r; b++) P.r2_thr[b] = p->r2_thr_host[b];
	for (int b = 0; b <= p->n_2; b++) P.thr2[b] = p->thr2_host[b];
	P.ncu = pl.g.ncu;
	P.ncv = pl.g.ncv;
	P.ncl = pl.g.ncl;
	P.inv_cu = pl.g.inv_cu;
	P.inv_cv = pl.g.inv_cv;
	P.inv_cl = pl.g.inv_cl;
	P.ku = pl.ku;
	P.kv = pl.kv;
	P.kl = pl.kl;
}

// Fixed-order reduction of the accumulator copies into the caller's output grids.
__global__ void k_finalize(const unsigned long long *__restrict__ cnt, const double *__restrict__ ddw,
						   const double *__restrict__ sp, const double *__restrict__ sc, int n_partials, int J, int nb,
						   int num_jk, mia_hist out) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	const size_t part = (size_t)2 * J * nb;
	unsigned long long t_cnt = 0;
	double t_ddw = 0.0, t_sp = 0.0, t_sc = 0.0;
	for (int k = 0; k < J; k++) {
		unsigned long long a_cnt = 0, b_cnt = 0;
		double a_ddw = 0.0, a_sp = 0.0, a_sc = 0.0, b_ddw = 0.0, b_sp = 0.0;
		for (int p = 0; p < n_partials; p++) {
			const size_t ia = p * part + (size_t)k * nb + b, ib = p * part + (size_t)(J + k) * nb + b;
			a_cnt += cnt[ia];
			a_ddw += ddw[ia];
			a_sp += sp[ia];
			a_sc += sc[ia];
			b_cnt += cnt[ib];
			b_ddw += ddw[ib];
			b_sp += sp[ib];
		}
		t_cnt += a_cnt;
		t_ddw += a_ddw;
		t_sp += a_sp;
		t_sc += a_sc;
		if (num_jk > 0) {
			if (out.dd_jk_count) out.dd_jk_count[(size_t)k * nb + b] = (int64_t)(a_cnt + b_cnt);
			if (out.dd_jk_w) out.dd_jk_w[(size_t)k * nb + b] = a_ddw + b_ddw;
			if (out.spd_jk) out.spd_jk[(size_t)k * nb + b] = a_sp + b_sp;
		}
	}
	out.dd_count[b] = (int64_t)t_cnt;
	out.dd_w[b] = t_ddw;
	out.spd[b] = t_sp;
	out.scd[b] = t_sc;
}

// Sum the accumulator copies (one per worker warp of the tiled kernel) into copy 0 in a fixed order, in two stages:
// RED_GROUPS consecutive ranges of copies are summed in parallel into a scratch block, then the group sums are added in
// group order.  Fixed association => reproducible bits.
__global__ void k_reduce_partials_stage1(const unsigned long long *__restrict__ cnt, const double *__restrict__ ddw,
										 const double *__restrict__ sp, const double *__restrict__ sc, int n_partials,
										 size_t n_el, unsigned long long *__restrict__ g_cnt, double *__restrict__ g_f) {
	const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	const int g = blockIdx.y;
	if (e >= n_el) return;
	const int per = (n_partials + RED_GROUPS - 1) / RED_GROUPS;
	const int p0 = g * per, p1 = (p0 + per < n_partials) ? p0 + per : n_partials;
	unsigned long long c = 0;
	double a = 0.0, b = 0.0, d = 0.0;
	for (int p = p0; p < p1; p++) {
		const size_t i = (size_t)p * n_el + e;
		c += cnt[i];
		a += ddw[i];
		b += sp[i];
		d += sc[i];
	}
	g_cnt[(size_t)g * n_el + e] = c;
	g_f[((size_t)g * 3 + 0) * n_el + e] = a;
	g_f[((size_t)g * 3 + 1) * n_el + e] = b;
	g_f[((size_t)g * 3 + 2) * n_el + e] = d;
}

__global__ void k_reduce_partials_stage2(unsigned long long *__restrict__ cnt, double *__restrict__ ddw,
										 double *__restrict__ sp, double *__restrict__ sc, size_t n_el,
										 const unsigned long long *__restrict__ g_cnt, const double *__restrict__ g_f) {
	const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (e >= n_el) return;
	unsigned long long c = 0;
	double a = 0.0, b = 0.0, d = 0.0;
	for (int g = 0; g < RED_GROUPS; g++) {
		c += g_cnt[(size_t)g * n_el + e];
		a += g_f[((size_t)g * 3 + 0) * n_el + e];
		b += g_f[((size_t)g * 3 + 1) * n_el + e];
		d += g_f[((size_t)g * 3 + 2) * n_el + e];
	}
	cnt[e] = c;
	ddw[e] = a;
	sp[e] = b;
	sc[e] = d;
}

// Fixed-order sum of the per-warp variance copies.
__global__ void k_reduce_var(const double *__restrict__ var, int n_partials, int nb, double *__restrict__ out) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	double s = 0.0;
	for (int p = 0; p < n_partials; p++) s += var[(size_t)p * nb + b];
	out[b] = s;
}

__global__ void k_copy_stats(const unsigned long long *in, uint64_t *out, unsigned long long kernel,
							 unsigned long long cells, unsigned long long tasks, unsigned long long launches) {
	if (threadIdx.x < 4) out[threadIdx.x] = in[threadIdx.x];
	if (threadIdx.x == 4) out[4] = kernel;
	if (threadIdx.x == 5) out[5] = cells;
	if (threadIdx.x == 6) out[6] = tasks ? tasks : in[6];
	if (threadIdx.x == 7) out[7] = launches;
}

__global__ void k_combine(const double *__restrict__ parts, int n_parts, int64_t n, double *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	double s = 0.0;
	for (int p = 0; p < n_parts; p++) s += parts[(size_t)p * n + i];
	out[i] = s;
}

}  // namespace

extern "C" {

const char *mia_strerror(int code) {
	switch (code) {
		case MIA_OK: return "ok";
		case MIA_ERR_ARG: return "bad argument";
		case MIA_ERR_WORKSPACE: return "workspace too small (see mia_workspace_bytes)";
		case MIA_ERR_RANGE: return "a coordinate lies outside [0, boxsize)";
		case MIA_ERR_WINDOW: return "tiled kernel: a pair fell outside its accumulation window (internal error)";
		case MIA_ERR_UNSUPPORTED: return "configuration not supported by the requested kernel";
		case MIA_ERR_UNSORTED: return "light-cone samples must be sorted by comoving distance (ascending, no NaN)";
		default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
	}
}

int mia_abi_version(void) { return MIA_ABI_VERSION; }

size_t mia_workspace_bytes(const mia_params *params, int64_t n_position, int64_t n_shape) {
	if (validate(params) != MIA_OK || n_position < 0 || n_shape < 0) return 0;
	Plan pl;
	if (make_plan(params, n_position, n_shape, pl) != MIA_OK) return 0;
	return pl.total + 256;
}

int mia_paircount(const mia_params *params, const mia_sample *D, const mia_sample *S, mia_shard shard,
				  const mia_hist *out, void *workspace, size_t workspace_bytes, void *stream) {
	int rc = validate(params);
	if (rc != MIA_OK) return rc;
	if (!D || !S || !out || D->n < 0 || S->n < 0) return MIA_ERR_ARG;
	if ((D->n > 0 && !D->pos) || (S->n > 0 && (!S->pos || !S->axis || !S->e))) return MIA_ERR_ARG;
	if (!out->dd_count || !out->dd_w || !out->spd || !out->scd) return MIA_ERR_ARG;
	if (params->num_jk > 0 && (!D->jk || !S->jk)) return MIA_ERR_ARG;
	if (params->variance && !out->var) return MIA_ERR_ARG;
	if (shard.count < 1 || shard.index < 0 || shard.index >= shard.count) return MIA_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int64_t nD = D->n, nS = S->n;

	Plan pl;
	rc = make_plan(params, nD, nS, pl);
	if (rc != MIA_OK) return rc;
	if (!workspace || workspace_bytes < pl.total) return MIA_ERR_WORKSPACE;
	unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
	if (ws + pl.total > (unsigned char *)workspace + workspace_bytes) return MIA_ERR_WORKSPACE;

	// same catalogue on both sides (auto-correlation): every unordered pair is visited once (mia_tiled_rppi2s.cuh)
	const bool alias = (D->pos == S->pos && D->weight == S->weight && D->jk == S->jk && nD == nS);
	pl.tiled.sym = (pl.kernel == MIA_KERNEL_TILED && pl.tiled.sym_ok && alias && pl.cand_bytes == (int)sizeof(CandSW)) ? 1 : 0;
	if (pl.tiled.sym && params->geometry == MIA_GEOM_RMU &&
		rmu_sym_chunk(D->weight == nullptr, pl.tiled.w_r * params->n_2) == 0)
		pl.tiled.sym = 0;  // weighted records + slots would leave one CTA per SM: the ordered kernel is faster
	if (pl.tiled.sym) pl.tiled.sym = rppi2s_cand_bytes(D->weight == nullptr);  // bytes per candidate record (48 / 64)
	DevParams P;
	fill_dev_params(params, pl, P);
	const int nl0 = (params->los == 0) ? 1 : 0, nl1 = (params->los == 2) ? 1 : 2, los = params->los;

	Cand *cand = (Cand *)(ws + pl.off_cand);
	int32_t *cand_jk = (int32_t *)(ws + pl.off_cand_jk);
	Prim *prim = (Prim *)(ws + pl.off_prim);
	int64_t *cell_start = (int64_t *)(ws + pl.off_cell_start);
	int64_t *prim_cell_start = (int64_t *)(ws + pl.off_prim_cell_start);
	SortScratch sc;
	sc.keys_in = (uint32_t *)(ws + pl.off_keys_in);
	sc.keys_out = (uint32_t *)(ws + pl.off_keys_out);
	sc.idx_in = (int32_t *)(ws + pl.off_idx_in);
	sc.idx_out = (int32_t *)(ws + pl.off_idx_out);
	sc.cub_tmp = ws + pl.off_cub;
	sc.cub_bytes = pl.cub_bytes;
	Accum A;
	A.cnt = (unsigned long long *)(ws + pl.off_cnt);
	A.ddw = (double *)(ws + pl.off_ddw);
	A.sp = (double *)(ws + pl.off_sp);
	A.sc = (double *)(ws + pl.off_sc);
	A.var = params->variance ? (double *)(ws + pl.off_var) : nullptr;
	A.stats = (unsigned long long *)(ws + pl.off_stats);
	A.rows = pl.rows;
	int *flags = (int *)(ws + pl.off_flags);

	// optional phase timing with CUDA events on the caller's stream (destroyed on every exit path)
	struct Events {
		cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
		~Events() {
			for (int i = 0; i < 4; i++)
				if (e[i]) cudaEventDestroy(e[i]);
		}
	} evs;
	cudaEvent_t *ev = evs.e;
	const bool timed = params->timings_host != nullptr;
	if (timed) {
		for (int i = 0; i < 4; i++) MIA_CUDA_CHECK(cudaEventCreate(&ev[i]));
		MIA_CUDA_CHECK(cudaEventRecord(ev[0], st));
	}
	unsigned long long n_launches = 0;

	// zero accumulators, stats and flags (contiguous region from off_cnt to off_tiled)
	MIA_CUDA_CHECK(cudaMemsetAsync(ws + pl.off_cnt, 0, pl.off_tiled - pl.off_cnt, st));

	const int T = 256;
	const int64_t ncell = pl.g.ncell();
	// ---- position sample -> sorted candidates + cell offsets ------------------------------------------------------
	rc = sort_by_cell(D->pos, D->jk, nD, nl0, nl1, los, pl.g, params->boxsize, sc, pl.key_bits, flags, st);
	if (rc) return rc;
	if (nD > 0 && pl.tiled.sym) {
		k_gather_cands<<<(unsigned)((nD + T - 1) / T), T, 0, st>>>(D->pos, D->weight, D->jk, S->axis, S->e, sc.idx_out, nD, nl0,
																	nl1, los, (void *)cand, cand_jk);
	} else if (nD > 0) {
		k_gather_cand<<<(unsigned)((nD + T - 1) / T), T, 0, st>>>(D->pos, D->weight, D->jk, sc.idx_out, nD, nl0, nl1, los,
																   cand, cand_jk);
	}
	k_cell_start<<<(unsigned)((nD + 1 + T - 1) / T), T, 0, st>>>(sc.keys_out, nD, pl.g.jk_rows, ncell, cell_start);
	MIA_CUDA_CHECK(cudaGetLastError());
	if (pl.kernel == MIA_KERNEL_TILED) {
		rc = tiled_prepare_candidates(pl.tiled, pl.g, P, sc.keys_out, cand, cand_jk, nD, cell_start, ws + pl.off_tiled, st);
		if (rc) return rc;
	}
	// ---- shape sample -> sorted primaries ---------------------------------------------------------------------------
	GridDims g_prim = pl.g;
	g_prim.sub = (pl.kernel == MIA_KERNEL_TILED) ? 2 : 1;
	g_prim.order = 0;
	if (pl.kernel == MIA_KERNEL_TILED && pl.tiled.ratio > 1) {  // (r, mu_r): the shape sample is sorted on coarser columns
		g_prim.ncu /= pl.tiled.ratio;
		g_prim.ncv /= pl.tiled.ratio;
		g_prim.inv_cu = g_prim.ncu / params->boxsize;
		g_prim.inv_cv = g_prim.ncv / params->boxsize;
	}
	rc = sort_by_cell(S->pos, S->jk, nS, nl0, nl1, los, g_prim, params->boxsize, sc, pl.key_bits, flags, st);
	if (rc) return rc;
	if (nS > 0) {
		k_gather_prim<<<(unsigned)((nS + T - 1) / T), T, 0, st>>>(S->pos, S->weight, S->jk, S->axis, S->e, sc.idx_out, nS,
																   nl0, nl1, los, prim);
	}
	k_cell_start<<<(unsigned)((nS + 1 + T - 1) / T), T, 0, st>>>(sc.keys_out, nS, pl.g.jk_rows, g_prim.nsorted_cells(),
																 prim_cell_start);
	MIA_CUDA_CHECK(cudaGetLastError());

	n_launches += (nD > 0 ? 2 : 0) + (nS > 0 ? 2 : 0) + 2;  // make_keys, gather (x2 samples) + 2 x cell_start
	Grid G;
	G.cand = cand;
	G.cand_jk = cand_jk;
	G.cell_start = cell_start;
	G.n_cand = nD;
	G.n_cell = ncell;

	// ---- shard of the (cell-sorted) shape sample ---------------------------------------------------------------------
	const int64_t s_begin = nS * shard.index / shard.count, s_end = nS * (shard.index + 1) / shard.count;
	unsigned long long n_tasks = 0;

	if (timed && pl.kernel == MIA_KERNEL_GENERAL) MIA_CUDA_CHECK(cudaEventRecord(ev[1], st));
	if (pl.kernel == MIA_KERNEL_GENERAL) {
		if (s_end > s_begin && nD > 0) {
			const int TB = 128;
			const unsigned blocks = (unsigned)((s_end - s_begin + TB - 1) / TB);
			const size_t smem = (size_t)pl.nb * (3 * sizeof(double) + sizeof(unsigned int));
			if (params->geometry == MIA_GEOM_RPPI) {
				MIA_CUDA_CHECK(cudaFuncSetAttribute(k_general<MIA_GEOM_RPPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
													(int)smem));
				k_general<MIA_GEOM_RPPI><<<blocks, TB, smem, st>>>(P, G, prim, s_begin, s_end, A);
			} else {
				MIA_CUDA_CHECK(cudaFuncSetAttribute(k_general<MIA_GEOM_RMU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
													(int)smem));
				k_general<MIA_GEOM_RMU><<<blocks, TB, smem, st>>>(P, G, prim, s_begin, s_end, A);
			}
			MIA_CUDA_CHECK(cudaGetLastError());
			n_tasks = blocks;
			n_launches += 1;
		}
		if (timed) MIA_CUDA_CHECK(cudaEventRecord(ev[2], st));
	} else {
		const bool unit_w = (D->weight == nullptr && S->weight == nullptr);
		rc = tiled_launch(pl.tiled, pl.g, g_prim, P, G, prim, prim_cell_start, nS, unit_w, shard, A, ws + pl.off_tiled, flags, st,
						  timed ? ev[1] : nullptr, timed ? ev[2] : nullptr);
		if (rc) return rc;
		n_launches += 1 /* cell_info */ + 2 /* col_chunks, fill_tasks */ + 1 /* pair kernel */ + 2 /* reduce_partials */ +
					  (params->geometry == MIA_GEOM_RMU ? 1 : 0) /* col_info */;
		if (pl.n_partials > 1) {
			const size_t n_el = (size_t)pl.rows * pl.nb;
			unsigned long long *g_cnt = (unsigned long long *)(ws + pl.off_red_cnt);
			double *g_f = (double *)(ws + pl.off_red_f);
			const dim3 grid1((unsigned)((n_el + 127) / 128), RED_GROUPS);
			k_reduce_partials_stage1<<<grid1, 128, 0, st>>>(A.cnt, A.ddw, A.sp, A.sc, pl.n_partials, n_el, g_cnt, g_f);
			k_reduce_partials_stage2<<<(unsigned)((n_el + 127) / 128), 128, 0, st>>>(A.cnt, A.ddw, A.sp, A.sc, n_el, g_cnt, g_f);
			MIA_CUDA_CHECK(cudaGetLastError());
		}
	}

	// ---- fixed-order final reduction ----------------------------------------------------------------------------------
	const int J = params->num_jk > 0 ? params->num_jk : 1;
	k_finalize<<<(pl.nb + 127) / 128, 128, 0, st>>>(A.cnt, A.ddw, A.sp, A.sc, 1, J, pl.nb, params->num_jk, *out);
	MIA_CUDA_CHECK(cudaGetLastError());
	if (params->variance) {
		k_reduce_var<<<(pl.nb + 127) / 128, 128, 0, st>>>(A.var, pl.n_partials, pl.nb, out->var);
		MIA_CUDA_CHECK(cudaGetLastError());
		n_launches += 1;
	}
	if (out->stats) {
		n_launches += 2;  // finalize + copy_stats
		k_copy_stats<<<1, 32, 0, st>>>(A.stats, out->stats,
									   (unsigned long long)(pl.tiled.sym ? MIA_KERNEL_TILED_SYM : pl.kernel), (unsigned long long)ncell,
									   n_tasks, n_launches);
		MIA_CUDA_CHECK(cudaGetLastError());
	}
	// range / window flags are checked synchronously: a wrong answer must never be returned silently
	int h_flags[8];
	MIA_CUDA_CHECK(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
	if (timed) MIA_CUDA_CHECK(cudaEventRecord(ev[3], st));
	MIA_CUDA_CHECK(cudaStreamSynchronize(st));
	if (timed) {
		float *t = params->timings_host;
		cudaEventElapsedTime(&t[0], ev[0], ev[1]);
		cudaEventElapsedTime(&t[1], ev[1], ev[2]);
		cudaEventElapsedTime(&t[2], ev[2], ev[3]);
		cudaEventElapsedTime(&t[3], ev[0], ev[3]);
	}
	if (h_flags[0]) return MIA_ERR_RANGE;
	if (h_flags[1]) return MIA_ERR_WINDOW;
	return MIA_OK;
}

int mia_paircount_host(const mia_params *params, const mia_sample *Dh, const mia_sample *Sh, mia_shard shard,
					   const mia_hist *out_h, int device) {
	int rc = validate(params);
	if (rc != MIA_OK) return rc;
	if (!Dh || !Sh || !out_h) return MIA_ERR_ARG;
	MIA_CUDA_CHECK(cudaSetDevice(device));
	const int64_t nD = Dh->n, nS = Sh->n;
	const bool same = (Dh->pos == Sh->pos && Dh->weight == Sh->weight && Dh->jk == Sh->jk && nD == nS);
	const int nb = params->n_r * params->n_2;
	const int J = params->num_jk;
	const size_t ws_bytes = mia_workspace_bytes(params, nD, nS);
	if (ws_bytes == 0) return MIA_ERR_ARG;

	// one device allocation, carved
	size_t o = 0;
	auto take = [&](size_t bytes) {
		size_t at = o;
		o = align_up(o + bytes);
		return at;
	};
	const size_t o_dpos = take(sizeof(double) * 3 * nD), o_dw = take(Dh->weight ? sizeof(double) * nD : 0),
				 o_djk = take(Dh->jk ? sizeof(int32_t) * nD : 0);
	const size_t o_spos = same ? o_dpos : take(sizeof(double) * 3 * nS);
	const size_t o_sw = same ? o_dw : take(Sh->weight ? sizeof(double) * nS : 0);
	const size_t o_sjk = same ? o_djk : take(Sh->jk ? sizeof(int32_t) * nS : 0);
	const size_t o_axis = take(sizeof(double) * 2 * nS), o_e = take(sizeof(double) * nS);
	const size_t o_cnt = take(sizeof(int64_t) * nb), o_ddw = take(sizeof(double) * nb), o_sp = take(sizeof(double) * nb),
				 o_sc = take(sizeof(double) * nb);
	const size_t o_jcnt = take(sizeof(int64_t) * (size_t)J * nb), o_jddw = take(sizeof(double) * (size_t)J * nb),
				 o_jsp = take(sizeof(double) * (size_t)J * nb);
	const size_t o_stats = take(sizeof(uint64_t) * 8);
	const size_t o_var = take(params->variance ? sizeof(double) * nb : 0);
	const size_t o_ws = take(ws_bytes);
	unsigned char *d = nullptr;
	MIA_CUDA_CHECK(cudaMalloc(&d, o + 256));
	cudaStream_t st;
	cudaError_t ce = cudaStreamCreate(&st);
	if (ce != cudaSuccess) {
		cudaFree(d);
		return (int)ce;
	}
#define H2D(off, src, bytes)                                                                         \
	if ((src) && (bytes) > 0 && rc == MIA_OK) {                                                      \
		cudaError_t _e = cudaMemcpyAsync(d + (off), (src), (bytes), cudaMemcpyHostToDevice, st);     \
		if (_e != cudaSuccess) rc = (int)_e;                                                         \
	}
	H2D(o_dpos, Dh->pos, sizeof(double) * 3 * nD);
	H2D(o_dw, Dh->weight, sizeof(double) * nD);
	H2D(o_djk, Dh->jk, sizeof(int32_t) * nD);
	if (!same) {
		H2D(o_spos, Sh->pos, sizeof(double) * 3 * nS);
		H2D(o_sw, Sh->weight, sizeof(double) * nS);
		H2D(o_sjk, Sh->jk, sizeof(int32_t) * nS);
	}
	H2D(o_axis, Sh->axis, sizeof(double) * 2 * nS);
	H2D(o_e, Sh->e, sizeof(double) * nS);
#undef H2D
	if (rc == MIA_OK) {
		mia_sample Dd = {nD, (const double *)(d + o_dpos), Dh->weight ? (const double *)(d + o_dw) : nullptr,
						 Dh->jk ? (const int32_t *)(d + o_djk) : nullptr, nullptr, nullptr};
		mia_sample Sd = {nS, (const double *)(d + o_spos), Sh->weight ? (const double *)(d + o_sw) : nullptr,
						 Sh->jk ? (const int32_t *)(d + o_sjk) : nullptr, (const double *)(d + o_axis),
						 (const double *)(d + o_e)};
		mia_hist od;
		od.dd_count = (int64_t *)(d + o_cnt);
		od.dd_w = (double *)(d + o_ddw);
		od.spd = (double *)(d + o_sp);
		od.scd = (double *)(d + o_sc);
		od.dd_jk_count = J > 0 ? (int64_t *)(d + o_jcnt) : nullptr;
		od.dd_jk_w = J > 0 ? (double *)(d + o_jddw) : nullptr;
		od.spd_jk = J > 0 ? (double *)(d + o_jsp) : nullptr;
		od.stats = (uint64_t *)(d + o_stats);
		od.var = params->variance ? (double *)(d + o_var) : nullptr;
		rc = mia_paircount(params, &Dd, &Sd, shard, &od, d + o_ws, ws_bytes, (void *)st);
	}
#define D2H(dst, off, bytes)                                                                         \
	if ((dst) && (bytes) > 0 && rc == MIA_OK) {                                                      \
		cudaError_t _e = cudaMemcpyAsync((dst), d + (off), (bytes), cudaMemcpyDeviceToHost, st);     \
		if (_e != cudaSuccess) rc = (int)_e;                                                         \
	}
	D2H(out_h->dd_count, o_cnt, sizeof(int64_t) * nb);
	D2H(out_h->dd_w, o_ddw, sizeof(double) * nb);
	D2H(out_h->spd, o_sp, sizeof(double) * nb);
	D2H(out_h->scd, o_sc, sizeof(double) * nb);
	D2H(out_h->dd_jk_count, o_jcnt, sizeof(int64_t) * (size_t)J * nb);
	D2H(out_h->dd_jk_w, o_jddw, sizeof(double) * (size_t)J * nb);
	D2H(out_h->spd_jk, o_jsp, sizeof(double) * (size_t)J * nb);
	D2H(out_h->stats, o_stats, sizeof(uint64_t) * 8);
	if (params->variance) D2H(out_h->var, o_var, sizeof(double) * nb);
#undef D2H
	cudaError_t se = cudaStreamSynchronize(st);
	if (rc == MIA_OK && se != cudaSuccess) rc = (int)se;
	cudaStreamDestroy(st);
	cudaFree(d);
	return rc;
}

int mia_combine_partials_f64(const double *parts, int32_t n_parts, int64_t n_values, double *out, void *stream) {
	if (!parts || !out || n_parts < 1 || n_values < 0) return MIA_ERR_ARG;
	if (n_values == 0) return MIA_OK;
	k_combine<<<(unsigned)((n_values + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parts, n_parts, n_values, out);
	MIA_CUDA_CHECK(cudaGetLastError());
	return MIA_OK;
}

// ---- light-cone brute pair loops (mia_lightcone.cuh) -------------------------------------------------------------------
static int lc_validate(const mia_lc_params *p, const mia_lc_sample *D, const mia_lc_sample *S, const mia_hist *out) {
	if (!p || p->abi_version != MIA_ABI_VERSION || !D || !S || !out) return MIA_ERR_ARG;
	if (p->n_r < 1 || p->n_r > MIA_MAX_BINS || p->n_2 < 1 || p->n_2 > MIA_MAX_BINS || p->num_patches < 0) return MIA_ERR_ARG;
	if (p->geometry != MIA_GEOM_RPPI && p->geometry != MIA_GEOM_RMU) return MIA_ERR_ARG;
	if (!p->r2_thr_host || !p->thr2_host || !(p->proj_scale > 0.0)) return MIA_ERR_ARG;
	if (D->n < 0 || S->n < 0 || D->n >= (1ll << 31) || S->n >= (1ll << 31)) return MIA_ERR_ARG;
	if (D->n > 0 && (!D->ra || !D->dec || !D->chi || !D->cosdec)) return MIA_ERR_ARG;
	if (S->n > 0 && (!S->ra || !S->dec || !S->chi)) return MIA_ERR_ARG;
	if (p->shapes && S->n > 0 && (!S->e1 || !S->e2)) return MIA_ERR_ARG;
	if (p->num_patches > 0 && ((D->n > 0 && !D->patch) || (S->n > 0 && !S->patch))) return MIA_ERR_ARG;
	if (!out->dd_count || !out->dd_w || !out->spd || !out->scd || !out->stats) return MIA_ERR_ARG;
	if (p->num_patches > 0 && (!out->dd_jk_count || !out->dd_jk_w || !out->spd_jk)) return MIA_ERR_ARG;
	return MIA_OK;
}

int mia_lightcone_paircount(const mia_lc_params *p, const mia_lc_sample *D, const mia_lc_sample *S, mia_shard shard,
							const mia_hist *out, void *stream) {
	int rc = lc_validate(p, D, S, out);
	if (rc != MIA_OK) return rc;
	if (shard.count < 1 || shard.index < 0 || shard.index >= shard.count) return MIA_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int nb = p->n_r * p->n_2;
	const size_t K = (size_t)p->num_patches;
	LcDev P;
	memset(&P, 0, sizeof(P));
	P.geom = p->geometry;
	P.n_r = p->n_r;
	P.n_2 = p->n_2;
	P.num_patches = p->num_patches;
	P.shapes = p->shapes ? 1 : 0;
	P.proj_scale = p->proj_scale;
	P.scaled = (p->proj_scale != 1.0) ? 1 : 0;
	P.rp2_cut = p->rp2_cut;
	for (int b = 0; b <= p->n_r; b++) P.r2_thr[b] = p->r2_thr_host[b];
	for (int b = 0; b <= p->n_2; b++) P.thr2[b] = p->thr2_host[b];
	P.reach = sqrt(P.r2_thr[p->n_r]) * (1.0 + 1e-9);
	P.cull_scale = (p->geometry == MIA_GEOM_RPPI) ? p->proj_scale : 1.0;  // the (r, mu_r) range test sees the unscaled dx, dy
	if (p->geometry == MIA_GEOM_RPPI) {
		P.win_lo = P.thr2[0];
		P.win_hi = P.thr2[p->n_2];
	} else {  // Pi^2 <= r^2 < r2_thr[n_r]
		const double rmax = sqrt(P.r2_thr[p->n_r]) * (1.0 + 1e-12);
		P.win_lo = -rmax;
		P.win_hi = rmax;
	}
	CudaEvents ev;
	const bool timed = p->timings_host != nullptr;
	if (timed) {
		rc = ev.create(4);
		if (rc) return rc;
		MIA_CUDA_CHECK(cudaEventRecord(ev.e[0], st));
	}
	MIA_CUDA_CHECK(cudaMemsetAsync(out->dd_count, 0, sizeof(int64_t) * nb, st));
	MIA_CUDA_CHECK(cudaMemsetAsync(out->dd_w, 0, sizeof(double) * nb, st));
	MIA_CUDA_CHECK(cudaMemsetAsync(out->spd, 0, sizeof(double) * nb, st));
	MIA_CUDA_CHECK(cudaMemsetAsync(out->scd, 0, sizeof(double) * nb, st));
	if (K) {
		MIA_CUDA_CHECK(cudaMemsetAsync(out->dd_jk_count, 0, sizeof(int64_t) * K * nb, st));
		MIA_CUDA_CHECK(cudaMemsetAsync(out->dd_jk_w, 0, sizeof(double) * K * nb, st));
		MIA_CUDA_CHECK(cudaMemsetAsync(out->spd_jk, 0, sizeof(double) * K * nb, st));
	}
	MIA_CUDA_CHECK(cudaMemsetAsync(out->stats, 0, sizeof(uint64_t) * 8, st));
	unsigned long long n_launch = 0;
	int *flag = reinterpret_cast<int *>(out->stats + 3);  // sortedness flag (stats[3], zeroed above)
	if (D->n > 0) {
		k_lc_check_sorted<<<(unsigned)((D->n + 255) / 256), 256, 0, st>>>(D->chi, D->n, flag);
		n_launch++;
	}
	if (S->n > 0) {
		k_lc_check_sorted<<<(unsigned)((S->n + 255) / 256), 256, 0, st>>>(S->chi, S->n, flag);
		n_launch++;
	}
	MIA_CUDA_CHECK(cudaGetLastError());
	int h_flag = 0;
	MIA_CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
	MIA_CUDA_CHECK(cudaStreamSynchronize(st));  // a wrong answer must never be returned silently
	if (h_flag) return MIA_ERR_UNSORTED;

	const int64_t p_begin = D->n * shard.index / shard.count, p_end = D->n * (shard.index + 1) / shard.count;
	if (timed) MIA_CUDA_CHECK(cudaEventRecord(ev.e[1], st));
	if (p_end > p_begin && S->n > 0) {
		LcSampleDev Dd = {D->n, D->ra, D->dec, D->chi, D->cosdec, D->weight, nullptr, nullptr, D->patch};
		LcSampleDev Sd = {S->n, S->ra, S->dec, S->chi, nullptr, S->weight, S->e1, S->e2, S->patch};
		LcOut O = {(unsigned long long *)out->dd_count, out->dd_w, out->spd, out->scd, (unsigned long long *)out->dd_jk_count,
				   out->dd_jk_w, out->spd_jk, (unsigned long long *)out->stats};
		const unsigned gx = (unsigned)((p_end - p_begin + LC_TP - 1) / LC_TP);
		int dev = 0, sms = 148;
		if (cudaGetDevice(&dev) == cudaSuccess) {
			int v = 0;
			if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
		}
		// segments of the chi window per position block: enough CTAs to fill the GPU a few times over
		long long n_seg = (8ll * sms + gx - 1) / gx;
		const long long max_tiles = (S->n + LC_TILE - 1) / LC_TILE;
		if (n_seg > max_tiles) n_seg = max_tiles;
		if (n_seg > 1024) n_seg = 1024;
		// a CTA counts its binned pairs per bin in 32 bits: at most 128 x 128 x (tiles per CTA) < 2^32
		const long long min_seg = (max_tiles + 200000 - 1) / 200000;
		if (n_seg < min_seg) n_seg = min_seg;
		if (n_seg < 1) n_seg = 1;
		const size_t smem = lightcone_smem_bytes(nb);
		const dim3 grid(gx, (unsigned)n_seg);
#define LC_LAUNCH(G, SH)                                                                                                       \
	do {                                                                                                                       \
		MIA_CUDA_CHECK(cudaFuncSetAttribute(k_lightcone<G, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
		k_lightcone<G, SH><<<grid, LC_TP, smem, st>>>(P, Dd, Sd, p_begin, p_end, (int)n_seg, O);                               \
	} while (0)
		if (p->geometry == MIA_GEOM_RPPI) {
			if (p->shapes) LC_LAUNCH(MIA_GEOM_RPPI, true);
			else LC_LAUNCH(MIA_GEOM_RPPI, false);
		} else {
			if (p->shapes) LC_LAUNCH(MIA_GEOM_RMU, true);
			else LC_LAUNCH(MIA_GEOM_RMU, false);
		}
#undef LC_LAUNCH
		MIA_CUDA_CHECK(cudaGetLastError());
		n_launch++;
	}
	if (timed) MIA_CUDA_CHECK(cudaEventRecord(ev.e[2], st));
	const unsigned long long tail[2] = {(unsigned long long)MIA_KERNEL_LIGHTCONE, n_launch};
	MIA_CUDA_CHECK(cudaMemcpyAsync(out->stats + 4, &tail[0], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	MIA_CUDA_CHECK(cudaMemcpyAsync(out->stats + 7, &tail[1], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	if (timed) MIA_CUDA_CHECK(cudaEventRecord(ev.e[3], st));
	MIA_CUDA_CHECK(cudaStreamSynchronize(st));
	if (timed) {
		float a = 0.f, b = 0.f;
		MIA_CUDA_CHECK(cudaEventElapsedTime(&a, ev.e[1], ev.e[2]));
		MIA_CUDA_CHECK(cudaEventElapsedTime(&b, ev.e[0], ev.e[3]));
		p->timings_host[0] = a;
		p->timings_host[1] = b;
	}
	return MIA_OK;
}

int mia_lightcone_paircount_host(const mia_lc_params *p, const mia_lc_sample *Dh, const mia_lc_sample *Sh, mia_shard shard,
								 const mia_hist *out_h, int device) {
	int rc = lc_validate(p, Dh, Sh, out_h);
	if (rc != MIA_OK) return rc;
	MIA_CUDA_CHECK(cudaSetDevice(device));
	const int nb = p->n_r * p->n_2;
	const size_t K = (size_t)p->num_patches;
	size_t o = 0;
	auto take = [&](size_t bytes) {
		size_t at = o;
		o = align_up(o + bytes);
		return at;
	};
	// inputs: up to 8 arrays per sample, then the outputs
	const double *srcD[7] = {Dh->ra, Dh->dec, Dh->chi, Dh->cosdec, Dh->weight, nullptr, nullptr};
	const double *srcS[7] = {Sh->ra, Sh->dec, Sh->chi, nullptr, Sh->weight, p->shapes ? Sh->e1 : nullptr, p->shapes ? Sh->e2 : nullptr};
	size_t offD[7], offS[7];
	for (int i = 0; i < 7; i++) offD[i] = take(srcD[i] ? sizeof(double) * Dh->n : 0);
	for (int i = 0; i < 7; i++) offS[i] = take(srcS[i] ? sizeof(double) * Sh->n : 0);
	const size_t o_pd = take(Dh->patch ? sizeof(int32_t) * Dh->n : 0), o_ps = take(Sh->patch ? sizeof(int32_t) * Sh->n : 0);
	const size_t o_cnt = take(sizeof(int64_t) * nb), o_ddw = take(sizeof(double) * nb), o_sp = take(sizeof(double) * nb),
				 o_sc = take(sizeof(double) * nb);
	const size_t o_jcnt = take(sizeof(int64_t) * K * nb), o_jddw = take(sizeof(double) * K * nb), o_jsp = take(sizeof(double) * K * nb);
	const size_t o_stats = take(sizeof(uint64_t) * 8);
	unsigned char *d = nullptr;
	MIA_CUDA_CHECK(cudaMalloc(&d, o + 256));
	cudaStream_t st;
	cudaError_t ce = cudaStreamCreate(&st);
	if (ce != cudaSuccess) {
		cudaFree(d);
		return (int)ce;
	}
	auto h2d = [&](size_t off, const void *src, size_t bytes) {
		if (src && bytes > 0 && rc == MIA_OK) {
			cudaError_t e_ = cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, st);
			if (e_ != cudaSuccess) rc = (int)e_;
		}
	};
	for (int i = 0; i < 7; i++) h2d(offD[i], srcD[i], sizeof(double) * Dh->n);
	for (int i = 0; i < 7; i++) h2d(offS[i], srcS[i], sizeof(double) * Sh->n);
	h2d(o_pd, Dh->patch, sizeof(int32_t) * Dh->n);
	h2d(o_ps, Sh->patch, sizeof(int32_t) * Sh->n);
	if (rc == MIA_OK) {
		auto dp = [&](const double *src, size_t off) { return src ? (const double *)(d + off) : (const double *)nullptr; };
		mia_lc_sample Dd = {Dh->n, dp(srcD[0], offD[0]), dp(srcD[1], offD[1]), dp(srcD[2], offD[2]), dp(srcD[3], offD[3]),
							dp(srcD[4], offD[4]), nullptr, nullptr, Dh->patch ? (const int32_t *)(d + o_pd) : nullptr};
		mia_lc_sample Sd = {Sh->n, dp(srcS[0], offS[0]), dp(srcS[1], offS[1]), dp(srcS[2], offS[2]), nullptr,
							dp(srcS[4], offS[4]), dp(srcS[5], offS[5]), dp(srcS[6], offS[6]),
							Sh->patch ? (const int32_t *)(d + o_ps) : nullptr};
		mia_hist od;
		memset(&od, 0, sizeof(od));
		od.dd_count = (int64_t *)(d + o_cnt);
		od.dd_w = (double *)(d + o_ddw);
		od.spd = (double *)(d + o_sp);
		od.scd = (double *)(d + o_sc);
		od.dd_jk_count = K ? (int64_t *)(d + o_jcnt) : nullptr;
		od.dd_jk_w = K ? (double *)(d + o_jddw) : nullptr;
		od.spd_jk = K ? (double *)(d + o_jsp) : nullptr;
		od.stats = (uint64_t *)(d + o_stats);
		rc = mia_lightcone_paircount(p, &Dd, &Sd, shard, &od, (void *)st);
	}
	auto d2h = [&](void *dst, size_t off, size_t bytes) {
		if (dst && bytes > 0 && rc == MIA_OK) {
			cudaError_t e_ = cudaMemcpyAsync(dst, d + off, bytes, cudaMemcpyDeviceToHost, st);
			if (e_ != cudaSuccess) rc = (int)e_;
		}
	};
	d2h(out_h->dd_count, o_cnt, sizeof(int64_t) * nb);
	d2h(out_h->dd_w, o_ddw, sizeof(double) * nb);
	d2h(out_h->spd, o_sp, sizeof(double) * nb);
	d2h(out_h->scd, o_sc, sizeof(double) * nb);
	d2h(out_h->dd_jk_count, o_jcnt, sizeof(int64_t) * K * nb);
	d2h(out_h->dd_jk_w, o_jddw, sizeof(double) * K * nb);
	d2h(out_h->spd_jk, o_jsp, sizeof(double) * K * nb);
	d2h(out_h->stats, o_stats, sizeof(uint64_t) * 8);
	cudaError_t se = cudaStreamSynchronize(st);
	if (rc == MIA_OK && se != cudaSuccess) rc = (int)se;
	cudaStreamDestroy(st);
	cudaFree(d);
	return rc;
}

}  // extern "C"

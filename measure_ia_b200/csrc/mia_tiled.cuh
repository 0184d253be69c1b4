// mia_tiled.cuh -- placeholder until the tiled kernel lands (plan_tiled returns false => general kernel).
#pragma once
#include "mia_common.cuh"
#include "mia_grid.cuh"

namespace mia {
struct TiledConfig { int n_partials; };
inline bool plan_tiled(const mia_params *, int64_t, int64_t, GridDims &, int &, int &, int &, TiledConfig &) { return false; }
inline size_t tiled_workspace_bytes(const TiledConfig &, const GridDims &, int64_t, int64_t) { return 0; }
inline int tiled_prepare_candidates(const TiledConfig &, const GridDims &, const DevParams &, const uint32_t *, const Cand *,
									int64_t, const int64_t *, void *, cudaStream_t) { return MIA_ERR_UNSUPPORTED; }
inline int tiled_launch(const TiledConfig &, const GridDims &, const DevParams &, const Grid &, const Prim *, const int64_t *,
						int64_t, int64_t, int64_t, const Accum &, void *, int *, unsigned long long *, cudaStream_t) {
	return MIA_ERR_UNSUPPORTED;
}
}  // namespace mia

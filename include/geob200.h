/* geob200 -- C ABI of the B200-native GeoTransformer registration hot path.
 *
 * Drop-in boundary.  Every entry point takes plain pointers and sizes; device
 * pointers unless the name ends in `_h`; `stream` is a cudaStream_t passed as void*.  All functions
 * return 0 on success and a negative code on failure, with a message available from geob200_last_error()
 * (the reference raises c10::Error -> RuntimeError through TORCH_CHECK, common/torch_helper.h:6-35; the
 * Python host layer turns a non-zero return into RuntimeError to keep that behaviour).
 *
 * Scratch memory is provided by the caller: each op has a *_workspace_bytes() query.
 * Index tables are int64 and sentinels equal the number of support rows, as in the reference.
 */
#ifndef GEOB200_H
#define GEOB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* geob200_last_error(void);

/* ---- collate -------------------------------------------------------------------------------------------- */

/* Replaces ext.grid_subsampling (reference geotransformer/extensions/pybind.cpp:13-17,
 * cpu/grid_subsampling/grid_subsampling.cpp:5-62): per-cloud voxel barycentres, bit-identical values AND order.
 * points (n_points,3) f32; lengths_h host int64[batch]; s_points must hold n_points*3 floats (upper bound);
 * s_lengths device int64[batch] receives the per-cloud counts (their sum = rows written). */
size_t geob200_grid_subsample_workspace_bytes(int64_t n_points, int64_t batch);
int geob200_grid_subsample(const float* points, int64_t n_points, const int64_t* lengths_h, int64_t batch, float voxel,
                           float* s_points, int64_t* s_lengths, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces ext.radius_neighbors (reference pybind.cpp:8-12, cpu/radius_neighbors/radius_neighbors.cpp:5-68)
 * fused with the column slice of modules/ops/radius_search.py:25-26.
 * For every query row: indices (offset by the cloud start) of the support points of the same batch element with
 * d2 < r*r, ascending by d2, first `width` of them, padded with the sentinel n_support.
 * max_count (device int32) receives the global maximum neighbour count (the reference's row width); a negative
 * value signals an unsupported density (> 16384 neighbours for one query).  counts (device int32[n_query]) and
 * out may be NULL (count-only pass when width == 0). */
size_t geob200_radius_search_workspace_bytes(int64_t n_query, int64_t n_support, int64_t batch);
int geob200_radius_search(const float* q_points, int64_t n_query, const float* s_points, int64_t n_support,
                          const int64_t* q_lengths_h, const int64_t* s_lengths_h, int64_t batch, float radius,
                          int64_t width, int64_t* out, int32_t* counts, int32_t* max_count, void* workspace,
                          size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOB200_H */

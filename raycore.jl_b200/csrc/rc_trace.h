// rc_trace.h — library-internal interface of the traversal kernels (rc_trace.cu) and analysis kernels (rc_analysis.cu)
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "rc_types.h"

#ifndef RC_TRACE_THREADS
#define RC_TRACE_THREADS 128
#endif

struct RcTraceLaunch {
    RcScene scene;
    const rc_ray *rays;  // device
    rc_hit *hits;        // device
    unsigned long long n;
    bool any, wide, count;
    bool zero_tmin = false;    // RC_IGNORE_TMIN (wide path only)
    bool watertight = false;   // RC_MODE_WATERTIGHT: the reference's watertight triangle test instead of Moeller-Trumbore
    unsigned long long *work;  // device work counter (zeroed by the launcher); with ovf_cap > 0: followed by the overflow list (work[1] = its
                               // length, work[2 ..] = ray indices, RcIoArrays::ovf_list), so that one memset clears both counters
    uint32_t ovf_cap = 0;
    RcCounters *counters;      // device, only with count
    uint32_t *overflow;        // device, incremented per ray whose traversal stack overflowed
    int max_blocks;            // persistent grid size (SMs x resident CTAs)
};

bool rc_launch_trace(cudaStream_t st, const RcTraceLaunch &L, std::string &err);
int rc_trace_max_blocks(int device);

// ---- analysis (rc_analysis.cu) ----
struct RcGridFrame {  // generate_ray_grid, src/kernels.jl:10-56, evaluated on the host in the reference's precision
    float gc[3], basis1[3], basis2[3], dir[3];
    float cell_w, cell_h;
    uint32_t grid;
};
bool rc_grid_frame(const float bounds[6], const float viewdir[3], uint32_t grid, RcGridFrame *out);
// fills rays (grid*grid) in Julia column-major cell order
void rc_launch_grid_rays(cudaStream_t st, const RcGridFrame &f, rc_ray *rays);
// hits_from_grid (:58-72): trace + hit point; points nullable; illum nullable (get_illumination :112-124, n_illum floats);
// centroid_acc nullable (3 doubles sum + 1 count as double, get_centroid :106-110)
void rc_launch_grid_trace(cudaStream_t st, const RcScene &sc, const RcGridFrame &f, rc_hit *hits, float *points, float *illum, uint32_t n_illum,
                          double *centroid_acc, uint32_t *overflow, int max_blocks);

struct RcFlatBlas {  // flat primitive array view: BLAS b covers flat positions [offset, offset + n)
    const RcTri *tris;
    uint32_t offset, n;
};
// view_factors! (:80-104): for every flat primitive whose metadata-1 lies in [row_base, row_base+n_rows) shoot rpt rays;
// out[(meta_src-1-row_base)*n_cols + meta_hit-1] += 1.  rays_out (nullable) receives the generated rays instead of tracing
// (ray (meta_src-1-row_base)*rpt + i).
void rc_launch_view_factors(cudaStream_t st, const RcScene &sc, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t rpt, unsigned long long seed,
                            uint32_t row_base, uint32_t n_rows, uint32_t n_cols, uint32_t *out, rc_ray *rays_out, unsigned long long *skipped,
                            uint32_t *overflow, int max_blocks, unsigned long long *work, const uint32_t *row_pos /* nullable, see rc_launch_vf_row_map */,
                            uint32_t row_stride /* owned rows: row_base + k * row_stride, k < n_rows */);
// row_pos[n_cols], info[2] = {duplicate metadata values, out-of-range metadata values}
void rc_launch_vf_row_map(cudaStream_t st, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t n_cols, uint32_t *row_pos, uint32_t *info);
void rc_launch_flat_metadata(cudaStream_t st, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t *out);

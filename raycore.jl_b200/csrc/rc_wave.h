// rc_wave.h — library-internal interface of the wavefront stage kernels (rc_wavefront.cu)
#pragma once
#include <cuda_runtime.h>

#include "rc_types.h"

struct RcCamera;
struct RcLights;

struct RcShadowSource {  // what a shadow ray is made from (all device pointers)
    const rc_ray *rays;               // primary rays
    const rc_hit *hits;               // their closest hits
    const rc_instance_desc *inst;     // instances[] (blas_index, inv_transform)
    const float *const *blas_normals; // per BLAS: 9 floats per primitive, indexed by hit.primitive_id
};

void rc_launch_gather_normals(cudaStream_t st, const RcTri *tris, uint32_t n, const float *d_in /* nullable: geometric */, float *d_out);
void rc_launch_primary_rays(cudaStream_t st, const RcCamera &cam, uint32_t width, uint32_t height, uint32_t n_samples, unsigned long long seed, rc_ray *d_rays);
void rc_launch_shadow_rays(cudaStream_t st, const RcShadowSource &s, const RcLights &lights, unsigned long long n_hits, rc_ray *d_out);
void rc_launch_test_shadow_rays(cudaStream_t st, const RcScene &sc, const rc_ray *d_rays, unsigned long long n, uint8_t *d_visible, uint32_t *overflow, int max_blocks,
                                unsigned long long *work);
void rc_launch_shadow_visibility(cudaStream_t st, const RcScene &sc, const RcShadowSource &s, const RcLights &lights, unsigned long long n_hits, uint8_t *d_visible,
                                 uint32_t *overflow, int max_blocks, unsigned long long *work);

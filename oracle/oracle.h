/* oracle/oracle.h — TEST INFRASTRUCTURE: C API of the CPU oracle (liboracle.so).
 *
 * The oracle is a CPU restatement of the reference's path-tracing hot path, evaluated on the
 * scene *object graph* (recursive lists, BVHs, ray-transforming instances, two boundary queries
 * per medium) rather than on the product's flattened buffers.  It is never linked into, imported
 * by, or called from the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Pinning status: the restated BVH builder is checked bit-for-bit against the reference's own
 * BVH.cu compiled from /root/reference (oracle/_ref/ref_bvh, CPU); the XORWOW integrator is
 * checked per pixel against the reference's own render_kernel compiled for sm_100a
 * (oracle/_ref/ref_render, fixtures under tests/golden/).  Quads, boxes, instances, media, lights,
 * image and Perlin textures do not exist in the reference: for those the oracle follows "Ray
 * Tracing: The Next Week" and parity is UNPINNED by the reference.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/rtb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

enum orc_rng_mode {
	ORC_RNG_PHILOX = 0,   /* the new path's spec: Philox4x32-10 keyed (seed,pixel) x (sample,bounce,stream) */
	ORC_RNG_XORWOW = 1    /* the reference's streams: curand XORWOW(seed + pixel), rejection samplers */
};

/* Parses an RTBS blob (include/rtb_scene_format.h). Returns NULL on error (orc_last_error()). */
orc_scene* orc_scene_load(const void* blob, size_t size);
void orc_scene_free(orc_scene* s);
const char* orc_last_error(void);

/* Renders samples [sample_begin, sample_end) of every pixel; ADDS into sum[w*h*4] (rgb sums, count)
 * and, if non-NULL, sum2[w*h*4] (sums of squares).  XORWOW mode requires sample_begin == 0.
 * rays_out (optional) receives the number of ray segments traced.  threads <= 0: all cores. */
int orc_render(const orc_scene* s, const rtb_camera* cam, uint32_t width, uint32_t height,
               uint32_t sample_begin, uint32_t sample_end, uint32_t max_depth, uint32_t seed,
               int rng_mode, int threads, float* sum, float* sum2, uint64_t* rays_out);

/* Closest-hit records for explicit rays (media skipped), same record layout as rtb_trace_rays. */
int orc_trace_rays(const orc_scene* s, const rtb_ray* rays, size_t n, rtb_hit* hits_out);

/* google_testing/test.cpp:87-106 recipe: brute-force closest sphere index over raw spheres
 * (center.xyz, radius) for a pinhole camera with u = x/(W-1)*2-1 pixel mapping. */
int orc_sphere_index_image(const float* spheres4, int n_spheres, const rtb_camera* cam, int width, int height, int32_t* out);

/* BVH_Handle::Factory restated (BVH.cu:166-383); same contract as rtb_bvh_build. */
int orc_bvh_build(const float* aabbs, int n, int builder, rtb_bvh_node* nodes_out, int* order_out, int* root_out);

/* Scalar kernels of the arithmetic spec, for unit tests. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_xorwow_uniforms(uint64_t seed, int n, float* out);
void orc_sincos2pi(float u, float* s, float* c);
float orc_logpos(float x);

#ifdef __cplusplus
}
#endif
#endif

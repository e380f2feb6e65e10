// rtb_kernels.cu — the wavefront path tracer's CUDA kernels (sm_100a).
//
//   generate    camera rays for one batch of (pixel, sample) paths          Renderer.cu:183-204, cu_Cameras.cuh:27-30,54-64,87-89
//   traverse    closest hit through the 64-byte two-box BVH nodes           BVH.cu:54-106, aabb.cuh:30-44, SphereHittable.cuh:15-33
//   shade       material scatter + texture evaluation + queue compaction    Renderer.cu:139-181, cu_materials.cuh:17-144, cu_Textures.cuh:9-40
//   accumulate  per-pixel sum of the batch's path contributions             Renderer.cu:204-206
//   resolve     mean -> clamp -> sqrt -> float4                             Renderer.cu:206-216
//
// Compiled with -fmad=false: every fused multiply-add below is explicit (see rtmath.h), so the
// CPU oracle reproduces the geometry bit for bit.  No tensor cores: nothing here is a dense
// contraction.  All kernels are persistent (grid = SMs x resident blocks) and pull work in
// warp- or block-sized chunks from a device counter, so the same launch sequence (and the same
// CUDA graph) serves every bounce regardless of how many paths are still alive.
#include "rtb_kernels.h"

#include <cfloat>

#include "rtb_types.h"
#include "rtmath.h"

namespace rtb {

using rt::v3;

#ifndef TRAVERSE_THREADS
#define TRAVERSE_THREADS 128
#endif
#ifndef SHADE_THREADS
#define SHADE_THREADS 128   // (B200, r2: 64 / 96 / 128 / 256 / 512 threads: shade 9.8 / 10.3 / 10.1 / 10.6 / 11.3 ms per batch, step 36.6-37.0 / 37.1 / 36.7 / 37.0 / 37.6:
#endif                      //  smaller blocks wait less at the compaction barriers, but interleave the survivors more finely for the next traverse)
#define STREAM_THREADS 256
#define DEEP_STACK_SIZE 64
#ifndef STACK_SIZE
#define STACK_SIZE 32
#endif
#define FULL_MASK 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// -DRTB_DEBUG_BOUNDS=1 (librtb200_debug.so): every index the kernels form from scene or queue data is checked against
// the size of what it indexes, and violations are counted per class instead of trapping (compute-sanitizer is not
// available on the pool this was developed on).  rtb_debug_bounds_report returns the counters; the release build
// compiles the checks away.
#ifndef RTB_DEBUG_BOUNDS
#define RTB_DEBUG_BOUNDS 0
#endif
#if RTB_DEBUG_BOUNDS
__device__ unsigned long long g_bounds_violations[RTB_BOUNDS_CLASSES];
__device__ unsigned long long g_bounds_checks;
#define RTB_CHECK(ok, cls) do { if (!(ok)) atomicAdd(&g_bounds_violations[cls], 1ull); } while (0)
#define RTB_COUNT_CHECKS(n) atomicAdd(&g_bounds_checks, (unsigned long long)(n))
#else
#define RTB_CHECK(ok, cls) do { } while (0)
#define RTB_COUNT_CHECKS(n) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// small device helpers

__device__ __forceinline__ v3 xyz(const float4& f) { return rt::mk(f.x, f.y, f.z); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
// Two consecutive float4 (32 bytes, 32-byte aligned).  sm_100 has 256-bit global loads (LDG.E.256): with
// -DRTB_LDG256=1 a 64-byte node is two requests instead of four.  Measured neutral on B200 (traverse 30.1 vs 29.9 ms
// per step, profiles/r1_experiments.md), so the default stays with 128-bit loads.
#ifndef RTB_LDG256
#define RTB_LDG256 0
#endif
__device__ __forceinline__ void ldg8(const float4* p, float4& a, float4& b) {
#if RTB_LDG256
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
	a = __ldg(p); b = __ldg(p + 1);
#endif
}

// Wavefront queue records are touched once per kernel.  -DRTB_STREAM_HINTS=1 marks those accesses streaming (evict-first,
// ld/st.global.cs) to keep them from pushing BVH nodes and primitive records out of L1 / L2.  Measured on B200 (round 2):
// traverse 28.4 vs 28.5 ms per step, shade 11.9 vs 11.4 ms - the queues are not what evicts the tree, and shade's stores
// do better write-back cached.  Off by default.
#ifndef RTB_STREAM_HINTS
#define RTB_STREAM_HINTS 0
#endif
#if RTB_STREAM_HINTS
__device__ __forceinline__ float4 ldq(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ int2 ldq(const int2* p) { return __ldcs(p); }
__device__ __forceinline__ void stq(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void stq(int2* p, int2 v) { __stcs(p, v); }
#else
__device__ __forceinline__ float4 ldq(const float4* p) { return *p; }
__device__ __forceinline__ int2 ldq(const int2* p) { return *p; }
__device__ __forceinline__ void stq(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void stq(int2* p, int2 v) { *p = v; }
#endif

// A ray record is 32 bytes, 32-byte aligned: one 256-bit access (LDG / STG.E.256 of sm_100) moves it, so a warp reading or
// writing 32 consecutive records touches 1 KB exactly once (two 128-bit accesses at a 32-byte stride ask L1 for every
// sector twice: shade measured 9 % slower that way).
__device__ __forceinline__ void ld_ray(const float4* rec, float4& o, float4& d) {
	asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(rec));
}
__device__ __forceinline__ void st_ray(float4* rec, const float4& o, const float4& d) {
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
	             :: "l"(rec), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w) : "memory");
}

// What changes from one rtb_render call to the next - the sample range and the seed - is read from device memory, so
// that the CUDA graph of a batch (whose kernel arguments are frozen at capture) serves every call on the same image size.
__device__ __forceinline__ BatchParams with_call_params(const BatchParams& in, const WaveView& wv) {
	BatchParams bp = in;
	bp.sample_begin = wv.call_params[0]; bp.sample_end = wv.call_params[1]; bp.seed = wv.call_params[2];
	return bp;
}

// Path id -> (global pixel index, absolute sample index).  Paths of a batch are laid out
// sample-major: id = local_sample * npix + local_pixel, local pixels row-major from row_begin.
__device__ __forceinline__ void path_pixel_sample(const BatchParams& bp, uint32_t batch, uint32_t path,
                                                  uint32_t& pixel, uint32_t& sample) {
	uint32_t sl = path / bp.npix;
	uint32_t pl = path - sl * bp.npix;
	pixel = bp.row_begin * bp.width + pl;
	sample = bp.sample_begin + batch * bp.samples_per_batch + sl;
}

__device__ __forceinline__ uint32_t batch_sample_count(const BatchParams& bp, uint32_t batch) {
	uint64_t s0 = (uint64_t)bp.sample_begin + (uint64_t)batch * bp.samples_per_batch;
	if (s0 >= bp.sample_end) return 0;
	uint64_t left = bp.sample_end - s0;
	return (uint32_t)(left < bp.samples_per_batch ? left : bp.samples_per_batch);
}

// ------------------------------------------------------------------------------------------------
// generate

__global__ void __launch_bounds__(STREAM_THREADS)
generate_kernel(BatchParams bp_in, rtb_camera cam, WaveView wv) {
	const BatchParams bp = with_call_params(bp_in, wv);
	const uint32_t batch = bp.batch_base + *wv.batch_index * bp.batch_stride;
	const uint32_t ns = batch_sample_count(bp, batch);
	const uint32_t n = ns * bp.npix;
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t stride = gridDim.x * blockDim.x;

	// Reset the per-bounce queue lengths and work counters of this batch.
	if (blockIdx.x == 0) {
		for (uint32_t i = threadIdx.x; i <= bp.max_depth; i += blockDim.x) wv.n_live[i] = (i == 0) ? n : 0u;
		for (uint32_t i = threadIdx.x; i < 2 * (bp.max_depth + 1); i += blockDim.x) wv.work[i] = 0u;
		for (uint32_t i = threadIdx.x; i <= bp.max_depth; i += blockDim.x) wv.n_tex[i] = 0u;
		if (threadIdx.x == 0) *wv.tail_from = 0xFFFFFFFFu;
	}

	const float px = 1.0f / (float)bp.width, py = 1.0f / (float)bp.height;   // pixel_size  Renderer.cu:188
	for (uint32_t path = tid; path < n; path += stride) {
		uint32_t pixel, sample;
		path_pixel_sample(bp, batch, path, pixel, sample);
		uint32_t y = pixel / bp.width, x = pixel - y * bp.width;
		rt::f4 r = rt::rng4(bp.seed, pixel, sample, RT_CAMERA_BOUNCE, rt::STREAM_SCATTER);
		// ndc = (vec2(x,y) + 0.5) * pixel_size * 2 - 1 ; jitter = point in unit disc * pixel_size   Renderer.cu:192,199
		float ndcx = fmaf(((float)x + 0.5f) * px, 2.0f, -1.0f);
		float ndcy = fmaf(((float)y + 0.5f) * py, 2.0f, -1.0f);
		float jx, jy; rt::unit_disc(r.x, r.y, jx, jy);
		float s = fmaf(jx, px, ndcx), t = fmaf(jy, py, ndcy);
		v3 o = rt::mk(cam.o[0], cam.o[1], cam.o[2]);
		v3 cu = rt::mk(cam.u[0], cam.u[1], cam.u[2]), cv = rt::mk(cam.v[0], cam.v[1], cam.v[2]), cw = rt::mk(cam.w[0], cam.w[1], cam.w[2]);
		v3 d; float time = 0.0f;
		if (cam.kind == RTB_CAM_DEFOCUS) {
			// DefocusBlurCamera::sample_ray  cu_Cameras.cuh:54-64
			rt::f4 l = rt::rng4(bp.seed, pixel, sample, RT_CAMERA_BOUNCE, rt::STREAM_LENS);
			float lx, ly; rt::unit_disc(l.x, l.y, lx, ly);
			v3 off = rt::mul(rt::mk(fmaf(cv.x, ly, cu.x * lx), fmaf(cv.y, ly, cu.y * lx), fmaf(cv.z, ly, cu.z * lx)), cam.lens_radius);
			v3 fwd = rt::mul(cw, cam.focus_dist);
			v3 hori = rt::mul(rt::mul(cu, cam.viewport_width), cam.focus_dist);
			v3 vert = rt::mul(rt::mul(cv, cam.viewport_height), cam.focus_dist);
			d = rt::sub(rt::madd(vert, t, rt::madd(hori, s, fwd)), off);
			o = rt::add(o, off);
			time = rt::mixf(cam.t0, cam.t1, r.z);
		} else {
			// PinholeCamera / MotionBlurCamera::sample_ray  cu_Cameras.cuh:27-30,87-89
			d = rt::madd(cv, t, rt::madd(cu, s, cw));
			if (cam.kind == RTB_CAM_MOTION) time = rt::mixf(cam.t0, cam.t1, r.z);
		}
		st_ray(wv.ray_od[0] + 2 * (size_t)path, make_float4(o.x, o.y, o.z, time), make_float4(d.x, d.y, d.z, __uint_as_float(path)));
		stq(wv.thr[0] + path, make_float4(1.0f, 1.0f, 1.0f, 0.0f));
		stq(wv.contrib + path, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
	}
}

// ------------------------------------------------------------------------------------------------
// primitive tests (shared by traverse and the hit-record hook)

// _sphere_closest_intersection  SphereHittable.cuh:15-33 (a = d.d hoisted per ray)
__device__ __forceinline__ float sphere_closest(v3 o, v3 d, float a, v3 c, float r) {
	v3 oc = rt::sub(o, c);
	float hb = rt::dot(d, oc);
	float cc = fmaf(-r, r, rt::dot(oc, oc));
	float disc = fmaf(hb, hb, -(a * cc));
	if (!(disc > 0.0f)) return FLT_MAX;
	float sq = sqrtf(disc);
	float t = (-hb - sq) / a;
	if (t < 0.0f) {
		t = (-hb + sq) / a;
		if (t < 0.0f) return FLT_MAX;
	}
	return t;
}

// Book quad::hit / triangle with the reference's t policy (t >= 0, strictly closer than the best).
__device__ __forceinline__ float planar_hit(v3 o, v3 d, const float4* pp, float4 q0, bool tri, float tbest) {
	float4 q1 = ldg4(pp + 1), q2, q3;
	ldg8(pp + 2, q2, q3);
	v3 N = rt::mk(q1.w, q2.w, q3.w);
	float denom = rt::dot(N, d);
	if (fabsf(denom) < 1e-8f) return FLT_MAX;
	float t = (q0.w - rt::dot(N, o)) / denom;
	if (!(t >= 0.0f) || !(t < tbest)) return FLT_MAX;
	v3 P = rt::madd(d, t, o);
	v3 planar = rt::sub(P, xyz(q0));
	v3 w = xyz(q3);
	float alpha = rt::dot(w, rt::cross(planar, xyz(q2)));
	float beta = rt::dot(w, rt::cross(xyz(q1), planar));
	if (tri) { if (alpha < 0.0f || beta < 0.0f || alpha + beta > 1.0f) return FLT_MAX; }
	else { if (alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f) return FLT_MAX; }
	return t;
}

// A book box() is six quads.  One BVH leaf holds a (min, max) record followed by the six ordinary quad records; a slab
// test on (min, max) only SELECTS which faces the ray can meet (the entry faces, then the exit faces, with a tolerance
// far above the rounding of either computation, in list order), and the hit itself is the exact quad test on those
// records - so the result is the one testing all six quads gives, and it is reported on the quad (`face_code`).
//   rec_stride = float4s from one record to the next (4, or 8 when every record is followed by its transform).
__device__ __forceinline__ float box_hit(v3 o, v3 d, float idx, float idy, float idz, float oix, float oiy, float oiz,
                                         const float4* pp, float4 q0, int rec_stride, float tbest, int& face) {
	const float4 q1 = ldg4(pp + 1);                         // q0 = (min.xyz, max.x), q1 = (max.y, max.z, -, -)
	const float ax = fmaf(q0.x, idx, oix), bx = fmaf(q0.w, idx, oix);
	const float ay = fmaf(q0.y, idy, oiy), by = fmaf(q1.x, idy, oiy);
	const float az = fmaf(q0.z, idz, oiz), bz = fmaf(q1.y, idz, oiz);
	const float t0x = fminf(ax, bx), t1x = fmaxf(ax, bx), t0y = fminf(ay, by), t1y = fmaxf(ay, by), t0z = fminf(az, bz), t1z = fmaxf(az, bz);
	const float tenter = fmaxf(fmaxf(t0x, t0y), t0z), texit = fminf(fminf(t1x, t1y), t1z);
	const float scale = fmaxf(fmaxf(fmaxf(fabsf(oix), fabsf(oiy)), fabsf(oiz)), fmaxf(fabsf(tenter), fabsf(texit)));
	const float tol = 2e-5f * scale;
	face = -1;
	if (tenter > texit + tol || texit < -tol || tenter - tol >= tbest) return FLT_MAX;
	// faces in list order: 0 front (z max), 1 right (x max), 2 back (z min), 3 left (x min), 4 top (y max), 5 bottom (y min)
	const int nfx = ax <= bx ? 3 : 1, nfy = ay <= by ? 5 : 4, nfz = az <= bz ? 2 : 0;
	unsigned entry = 0, exits = 0;
	if (t0x >= tenter - tol) entry |= 1u << nfx;
	if (t0y >= tenter - tol) entry |= 1u << nfy;
	if (t0z >= tenter - tol) entry |= 1u << nfz;
	if (t1x <= texit + tol) exits |= 1u << (4 - nfx);       // the opposite face: 3 <-> 1
	if (t1y <= texit + tol) exits |= 1u << (9 - nfy);       // 5 <-> 4
	if (t1z <= texit + tol) exits |= 1u << (2 - nfz);       // 2 <-> 0
	unsigned mask = tenter >= -tol ? entry : 0u;
	if (texit - tenter <= tol) mask |= exits;                // paper-thin along the ray: the order of entry and exit is not reliable
	float best = FLT_MAX;
	for (int pass = 0; pass < 2; ++pass) {
		for (unsigned m = mask; m; m &= m - 1) {
			const int f = __ffs(m) - 1;
			const float4* fp = pp + rec_stride * (1 + f);
			const float t = planar_hit(o, d, fp, ldg4(fp), false, fminf(tbest, best));
			if (t < best) { best = t; face = f; }
		}
		if (best < FLT_MAX) break;
		mask = exits & ~mask;
	}
	return best;
}

// Free-flight uniforms of the media a ray meets.  Medium m of a path segment uses component (m & 3)
// of Philox(seed, pixel; sample, bounce, STREAM_MEDIUM0 + (m >> 2)); the block is generated lazily,
// only when some medium's boundary interval survives the geometric rejections, and shared by up to
// four media.
struct MediumRng {
	uint32_t seed, path, bounce;
	uint32_t npix, pixel0, sample0;   // batch layout (warp-uniform): the path's pixel and sample are only worked out
	                                  // (an integer division) when a block of uniforms is actually drawn
	uint32_t block;    // cached block index, 0xFFFFFFFF = none
	rt::f4 u;
	__device__ __forceinline__ float get(uint32_t m) {
		const uint32_t b = m >> 2;
		if (b != block) {
			const uint32_t sl = path / npix, pl = path - sl * npix;      // path_pixel_sample
			u = rt::rng4(seed, pixel0 + pl, sample0 + sl, bounce, rt::STREAM_MEDIUM0 + b); block = b;
		}
		const uint32_t c = m & 3u;
		return c == 0 ? u.x : (c == 1 ? u.y : (c == 2 ? u.z : u.w));
	}
};
__device__ __forceinline__ MediumRng make_medium_rng(const BatchParams& bp, uint32_t batch, uint32_t path, uint32_t bounce) {
	MediumRng r; r.seed = bp.seed; r.path = path; r.bounce = bounce; r.block = 0xFFFFFFFFu;
	r.npix = bp.npix; r.pixel0 = bp.row_begin * bp.width; r.sample0 = bp.sample_begin + batch * bp.samples_per_batch;
	r.u.x = r.u.y = r.u.z = r.u.w = 0.0f; return r;
}

// constant_medium::hit (book) over a convex boundary interval [t1,t2] found for any-sign t.
__device__ __forceinline__ float medium_sample(float t1, float t2, float a, float neg_inv_density, MediumRng& mr, uint32_t medium, float tbest) {
	if (!(t2 > t1 + 0.0001f)) return FLT_MAX;     // second boundary query: interval(t1 + 0.0001, inf)
	if (t1 < 0.0f) t1 = 0.0f;                     // ray_t.min (the reference accepts t >= 0)
	if (t2 > tbest) t2 = tbest;                   // ray_t.max = closest so far
	if (t1 >= t2) return FLT_MAX;
	float len = sqrtf(a);
	float dist_inside = (t2 - t1) * len;
	float hit_distance = neg_inv_density * rt::logpos(mr.get(medium));
	if (hit_distance > dist_inside) return FLT_MAX;
	return t1 + hit_distance / len;
}

__device__ __forceinline__ float medium_sphere_hit(v3 o, v3 d, float a, float4 q0, float nid, MediumRng& mr, uint32_t medium, float tbest) {
	v3 oc = rt::sub(o, xyz(q0));
	float hb = rt::dot(d, oc);
	float cc = fmaf(-q0.w, q0.w, rt::dot(oc, oc));
	float disc = fmaf(hb, hb, -(a * cc));
	if (!(disc > 0.0f)) return FLT_MAX;
	float sq = sqrtf(disc);
	return medium_sample((-hb - sq) / a, (-hb + sq) / a, a, nid, mr, medium, tbest);
}

__device__ __forceinline__ float medium_box_hit(v3 o, v3 d, float a, float4 q0, float4 q1, float4 q2, MediumRng& mr, uint32_t medium, float tbest) {
	// world -> object: translate back, rotate by -theta about y (book translate::hit / rotate_y::hit)
	float cs = q0.w, sn = q1.w;
	v3 ot = rt::sub(o, xyz(q2));
	v3 oo = rt::mk(fmaf(cs, ot.x, -(sn * ot.z)), ot.y, fmaf(sn, ot.x, cs * ot.z));
	v3 dd = rt::mk(fmaf(cs, d.x, -(sn * d.z)), d.y, fmaf(sn, d.x, cs * d.z));
	float tn = -FLT_MAX, tf = FLT_MAX;
	const float bmin[3] = {q0.x, q0.y, q0.z}, bmax[3] = {q1.x, q1.y, q1.z};
	const float oc[3] = {oo.x, oo.y, oo.z}, dc[3] = {dd.x, dd.y, dd.z};
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		if (dc[k] == 0.0f) { if (oc[k] < bmin[k] || oc[k] > bmax[k]) return FLT_MAX; continue; }
		float ta = (bmin[k] - oc[k]) / dc[k], tb = (bmax[k] - oc[k]) / dc[k];
		float lo = ta < tb ? ta : tb, hi = ta < tb ? tb : ta;
		tn = lo > tn ? lo : tn; tf = hi < tf ? hi : tf;
	}
	if (!(tn < tf)) return FLT_MAX;
	return medium_sample(tn, tf, a, q2.w, mr, medium, tbest);
}

// Instances.  T = (cos, sin, off.x, off.y), (off.z, -, -, -) with world = R_y(theta) * object + off.
__device__ __forceinline__ int xf_offset(int base) { return base == PRIM_SPHERE ? 1 : (base == PRIM_MOVING_SPHERE ? 2 : 4); }   // quads, triangles, boxes: the next record

// world ray -> object ray: origin - offset (book translate::hit), then rotate by -theta (book rotate_y::hit)
__device__ __forceinline__ void xf_ray(const float4* tp, v3 o, v3 d, v3& oo, v3& dd) {
	const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1);
	const float cs = t0.x, sn = t0.y;
	const v3 ot = rt::sub(o, rt::mk(t0.z, t0.w, t1.x));
	if (t1.y == 0.0f) { oo = ot; dd = d; return; }   // translate only: the book's translate::hit leaves the direction alone (signed zeros included)
	oo = rt::mk(fmaf(cs, ot.x, -(sn * ot.z)), ot.y, fmaf(sn, ot.x, cs * ot.z));
	dd = rt::mk(fmaf(cs, d.x, -(sn * d.z)), d.y, fmaf(sn, d.x, cs * d.z));
}
// object vector -> world (book rotate_y::hit, normal back-rotation)
__device__ __forceinline__ v3 xf_vec_to_world(const float4* tp, v3 v) {
	const float4 t0 = ldg4(tp);
	const float cs = t0.x, sn = t0.y;
	if (ldg4(tp + 1).y == 0.0f) return v;   // translate only
	return rt::mk(fmaf(cs, v.x, sn * v.z), v.y, fmaf(-sn, v.x, cs * v.z));
}

// ------------------------------------------------------------------------------------------------
// traverse


// One primitive against one ray: the t of the hit if it is closer than tbest, else FLT_MAX.
// `hit_code` is what a hit is reported on: the leaf itself, or for a box the quad record of the face that was hit.
// (idx.. oiz: the ray's slab-test reciprocals, used to pick the faces of a box.)
struct RaySlab { float idx, idy, idz, oix, oiy, oiz; };

// Per-ray reciprocal of the slab tests.  A component that is (nearly) zero is nudged to +-1e-20 with the sign of the
// component - signed zeros included, and every later sign decision is taken from the nudged value - so 0 * inf never appears.
__device__ __forceinline__ float slab_dir(float d) { return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; }

#ifndef RTB_UNIFIED_LEAF
#define RTB_UNIFIED_LEAF 1
#endif
template <bool MEDIA>
__device__ __forceinline__ float leaf_test(const SceneView& sv, int code, v3 o, v3 d, float a, float time, MediumRng& mr, float tbest,
                                           const RaySlab& rs, int& hit_code) {
	hit_code = code;
	const int type = code & 15;
	RTB_CHECK(code >= 0 && (code >> RTB_LEAF_TYPE_BITS) < sv.n_prims, RTB_BOUNDS_PRIM);
	const float4* pp = sv.prims + 4 * (size_t)(code >> RTB_LEAF_TYPE_BITS);
	const float4 q0 = ldg4(pp);
	float t = FLT_MAX;
#if RTB_UNIFIED_LEAF
	// The lanes of a warp meet different flavours of the same shape (sphere / moving sphere / instanced sphere; box /
	// instanced box): the instance transform and the moving centre are short predicated preludes, and the long part -
	// the quadratic, the face tests - is ONE code path per shape that all those lanes run together.
	const int base = type & 7;
	if (MEDIA && (type == PRIM_MEDIUM_SPHERE || type == PRIM_MEDIUM_BOX)) {
		const float4 q1 = ldg4(pp + 1);
		if (type == PRIM_MEDIUM_SPHERE) return medium_sphere_hit(o, d, a, q0, q1.x, mr, __float_as_uint(q1.y), tbest);
		const float4 q2 = ldg4(pp + 2), q3 = ldg4(pp + 3);
		return medium_box_hit(o, d, a, q0, q1, q2, mr, __float_as_uint(q3.x), tbest);
	}
	const bool xf = (type & PRIM_XF) != 0;
	v3 oo = o, dd = d; float a2 = a;
	if (xf) {   // instance: take the ray into the primitive's frame (book translate::hit, rotate_y::hit)
		xf_ray(pp + xf_offset(base), o, d, oo, dd);
		a2 = rt::dot(dd, dd);
	}
	if (base <= PRIM_MOVING_SPHERE) {
		v3 c = xyz(q0);
		if (base == PRIM_MOVING_SPHERE) c = rt::mix(c, xyz(ldg4(pp + 1)), time);   // center = mix(center0, center1, ray.time)   SphereHittable.cu:92
		t = sphere_closest(oo, dd, a2, c, q0.w);
	} else if (base == PRIM_BOX) {
		// the box in its own frame; when instanced, every record (the box's and each face's) is followed by the transform
		RaySlab bs = rs;
		if (xf) {
			bs.idx = __frcp_rn(slab_dir(dd.x)); bs.idy = __frcp_rn(slab_dir(dd.y)); bs.idz = __frcp_rn(slab_dir(dd.z));
			bs.oix = -(oo.x * bs.idx); bs.oiy = -(oo.y * bs.idy); bs.oiz = -(oo.z * bs.idz);
		}
		const int stride = xf ? 8 : 4;
		int face;
		t = box_hit(oo, dd, bs.idx, bs.idy, bs.idz, bs.oix, bs.oiy, bs.oiz, pp, q0, stride, tbest, face);
		if (face >= 0) hit_code = (((code >> RTB_LEAF_TYPE_BITS) + (stride >> 2) * (1 + face)) << RTB_LEAF_TYPE_BITS) | PRIM_QUAD | (type & PRIM_XF);
	} else {
		t = planar_hit(oo, dd, pp, q0, base == PRIM_TRIANGLE, tbest);
	}
#else
	if (type == PRIM_SPHERE) {
		t = sphere_closest(o, d, a, xyz(q0), q0.w);
	} else if (type == PRIM_MOVING_SPHERE) {
		// center = mix(center0, center1, ray.time)   SphereHittable.cu:92
		const float4 q1 = ldg4(pp + 1);
		t = sphere_closest(o, d, a, rt::mix(xyz(q0), xyz(q1), time), q0.w);
	} else if (type == PRIM_QUAD || type == PRIM_TRIANGLE) {
		t = planar_hit(o, d, pp, q0, type == PRIM_TRIANGLE, tbest);
	} else if (type == PRIM_BOX) {
		int face;
		t = box_hit(o, d, rs.idx, rs.idy, rs.idz, rs.oix, rs.oiy, rs.oiz, pp, q0, 4, tbest, face);
		if (face >= 0) hit_code = (((code >> RTB_LEAF_TYPE_BITS) + 1 + face) << RTB_LEAF_TYPE_BITS) | PRIM_QUAD;
	} else if (type & PRIM_XF) {
		// instance: take the ray into the primitive's frame (book translate::hit, rotate_y::hit)
		const int base = type & 7;
		v3 oo, dd;
		xf_ray(pp + xf_offset(base), o, d, oo, dd);
		const float a2 = rt::dot(dd, dd);
		if (base == PRIM_SPHERE) t = sphere_closest(oo, dd, a2, xyz(q0), q0.w);
		else if (base == PRIM_MOVING_SPHERE) t = sphere_closest(oo, dd, a2, rt::mix(xyz(q0), xyz(ldg4(pp + 1)), time), q0.w);
		else if (base == PRIM_BOX) {
			// the box in its own frame; every record (the box's and each face's) is followed by the transform
			const float jx = __frcp_rn(slab_dir(dd.x)), jy = __frcp_rn(slab_dir(dd.y)), jz = __frcp_rn(slab_dir(dd.z));
			int face;
			t = box_hit(oo, dd, jx, jy, jz, -(oo.x * jx), -(oo.y * jy), -(oo.z * jz), pp, q0, 8, tbest, face);
			if (face >= 0) hit_code = (((code >> RTB_LEAF_TYPE_BITS) + 2 * (1 + face)) << RTB_LEAF_TYPE_BITS) | PRIM_QUAD | PRIM_XF;
		}
		else t = planar_hit(oo, dd, pp, q0, base == PRIM_TRIANGLE, tbest);
	} else if (MEDIA) {
		const float4 q1 = ldg4(pp + 1);
		if (type == PRIM_MEDIUM_SPHERE) {
			t = medium_sphere_hit(o, d, a, q0, q1.x, mr, __float_as_uint(q1.y), tbest);
		} else {
			const float4 q2 = ldg4(pp + 2), q3 = ldg4(pp + 3);
			t = medium_box_hit(o, d, a, q0, q1, q2, mr, __float_as_uint(q3.x), tbest);
		}
	}
#endif
	return t;
}

#ifndef TRAV_WHILE_WHILE
#define TRAV_WHILE_WHILE 1
#endif
#ifndef RTB_ROBUST_SLAB
#define RTB_ROBUST_SLAB 1
#endif
#ifndef RTB_REG_STACK
#define RTB_REG_STACK 0
#endif
#define TRAV_END ((int)0x80000000)

// MEDIA: 0 = the scene has no media; 1 = all of them are in the pre-test list (the walk itself never meets
// one: no RNG state is carried through the loop); 2 = more than the list holds, the rest are BVH leaves.
template <int MEDIA, bool STATS = false>
__device__ __forceinline__ void trace_ray(const SceneView& sv, v3 o, v3 d, float time, MediumRng& mr,
                                          int* __restrict__ stack, int stack_cap, float& tbest_out, int& code_out, int* stats_out = nullptr) {
	const float a = rt::dot(d, d);
	float tbest = FLT_MAX;
	int best = -1;
	if (MEDIA) {   // media first: a scatter inside a medium bounds the BVH walk tightly
		for (int k = 0; k < sv.n_pre; ++k) {
			const int code = __ldg(sv.pre_list + k);
			int hit_code;
			const float t = leaf_test<true>(sv, code, o, d, a, time, mr, tbest, RaySlab{}, hit_code);   // (media only: no slab data needed)
			if (t < tbest) { tbest = t; best = hit_code; }
		}
		if (sv.bvh_empty) { tbest_out = tbest; code_out = best; return; }
	}
	// Slab test with a per-ray reciprocal (slab_dir: zero components, signed zeros included, are nudged).
	const float gx = slab_dir(d.x), gy = slab_dir(d.y), gz = slab_dir(d.z);
	const float idx = __frcp_rn(gx), idy = __frcp_rn(gy), idz = __frcp_rn(gz);
	const float oix = -(o.x * idx), oiy = -(o.y * idy), oiz = -(o.z * idz);
	// Near / far slab planes are picked by the direction signs with FMAs instead of min/max pairs:
	// t(min plane) = min * (1/d) - o/d, and the max plane adds ext * (1/d) on the side the sign says.
	// (ncu: the min/max version kept the ALU pipe 66 % busy with the FMA pipe at 20 %.)
	// The sign is that of the value the reciprocal was taken of, so a -0.0 component picks the planes its -1e20 reciprocal needs.
	const float nx = gx < 0.0f ? idx : 0.0f, fx = gx < 0.0f ? 0.0f : idx;
	const float ny = gy < 0.0f ? idy : 0.0f, fy = gy < 0.0f ? 0.0f : idy;
	const float nz = gz < 0.0f ? idz : 0.0f, fz = gz < 0.0f ? 0.0f : idz;
#if RTB_ROBUST_SLAB
	// The cull must never lose a box the exact primitive test would hit (aabb::intersects computes (min - o) / d, which
	// cancels exactly; plane * (1/d) - o/d does not): the rounding of o/d is up to 2^-24 |o/d| per plane, the reciprocal and
	// the two fused multiply-adds add a few 2^-24 of t.  The far distance is widened by both bounds before it is compared -
	// one FMA per box; a box thinner than the rounding (a 1e-4 quad seen from 10,000 units) is then always entered.
	const float slab_abs = 2.4e-7f * fmaxf(fmaxf(fabsf(oix), fabsf(oiy)), fabsf(oiz));   // 2 x 2^-23 max |o/d|
	const float slab_rel = 1.0f + 9.6e-7f;                                                // 1 + 8 x 2^-23
#endif

#if RTB_NODE_PAIRED
	const float2 id_xy = make_float2(idx, idy), oi_xy = make_float2(oix, oiy), n_xy = make_float2(nx, ny), f_xy = make_float2(fx, fy);
	const float2 id_zz = make_float2(idz, idz), oi_zz = make_float2(oiz, oiz), n_zz = make_float2(nz, nz), f_zz = make_float2(fz, fz);
#endif
	int cur = sv.root_ref;
	int sp = 0;
	// -DRTB_REG_STACK=1 (experiment): the two youngest stack entries live in registers and only older ones spill to shared
	// memory - the "short stack in registers" of the brief.  Measured on B200 (round 2): see profiles/r2_experiments.md.
#if RTB_REG_STACK
	int top0 = 0, top1 = 0;   // top0 = youngest
#define STACK_PUSH(v) do { if (sp >= 2) stack[(sp - 2) * TRAVERSE_THREADS] = top1; top1 = top0; top0 = (v); ++sp; } while (0)
#define STACK_POP(dst) do { dst = top0; top0 = top1; --sp; if (sp >= 2) top1 = stack[(sp - 2) * TRAVERSE_THREADS]; } while (0)
#else
#define STACK_PUSH(v) do { stack[sp * TRAVERSE_THREADS] = (v); ++sp; } while (0)
#define STACK_POP(dst) do { --sp; dst = stack[sp * TRAVERSE_THREADS]; } while (0)
#endif
	int n_inner = 0, n_leaf = 0;   // STATS only
	// "while-while" walk: lanes first descend inner nodes together, then test their leaf primitive
	// together (a leaf reference is negative; TRAV_END marks an exhausted walk).
	for (;;) {
		while (cur >= 0) {
			if (STATS) ++n_inner;
			RTB_CHECK(cur < sv.n_nodes, RTB_BOUNDS_NODE);
			const float4* np = sv.nodes + 4 * (size_t)cur;
			float4 n0, n1, n2, n3f;
			ldg8(np, n0, n1); ldg8(np + 2, n2, n3f);
			const int2 n3 = make_int2(__float_as_int(n3f.x), __float_as_int(n3f.y));
#if RTB_NODE_PAIRED
			// the same 18 fused multiply-adds, two per instruction (FFMA2): x and y of one child share an instruction,
			// z of the two children share one; every component is an IEEE fma, so the results are the bits of the scalar form
			const float2 l_xy = __ffma2_rn(make_float2(n0.x, n0.y), id_xy, oi_xy);
			const float2 r_xy = __ffma2_rn(make_float2(n1.x, n1.y), id_xy, oi_xy);
			const float2 lr_z = __ffma2_rn(make_float2(n2.x, n2.y), id_zz, oi_zz);
			const float2 ln_xy = __ffma2_rn(make_float2(n0.z, n0.w), n_xy, l_xy), lf_xy = __ffma2_rn(make_float2(n0.z, n0.w), f_xy, l_xy);
			const float2 rn_xy = __ffma2_rn(make_float2(n1.z, n1.w), n_xy, r_xy), rf_xy = __ffma2_rn(make_float2(n1.z, n1.w), f_xy, r_xy);
			const float2 lrn_z = __ffma2_rn(make_float2(n2.z, n2.w), n_zz, lr_z), lrf_z = __ffma2_rn(make_float2(n2.z, n2.w), f_zz, lr_z);
			const float ltmin = fmaxf(fmaxf(ln_xy.x, ln_xy.y), lrn_z.x);
			const float ltmax = fminf(fminf(lf_xy.x, lf_xy.y), lrf_z.x);
			const float rtmin = fmaxf(fmaxf(rn_xy.x, rn_xy.y), lrn_z.y);
			const float rtmax = fminf(fminf(rf_xy.x, rf_xy.y), lrf_z.y);
#else
			const float lx = fmaf(n0.x, idx, oix), ly = fmaf(n0.y, idy, oiy), lz = fmaf(n0.z, idz, oiz);
			const float rx = fmaf(n1.z, idx, oix), ry = fmaf(n1.w, idy, oiy), rz = fmaf(n2.x, idz, oiz);
			const float ltmin = fmaxf(fmaxf(fmaf(n0.w, nx, lx), fmaf(n1.x, ny, ly)), fmaf(n1.y, nz, lz));
			const float ltmax = fminf(fminf(fmaf(n0.w, fx, lx), fmaf(n1.x, fy, ly)), fmaf(n1.y, fz, lz));
			const float rtmin = fmaxf(fmaxf(fmaf(n2.y, nx, rx), fmaf(n2.z, ny, ry)), fmaf(n2.w, nz, rz));
			const float rtmax = fminf(fminf(fmaf(n2.y, fx, rx), fmaf(n2.z, fy, ry)), fmaf(n2.w, fz, rz));
#endif
			// aabb::intersects: tmin <= tmax && tmin < ray_max && tmax > 0   aabb.cuh:41
#if RTB_ROBUST_SLAB
#if RTB_NODE_PAIRED
			const float2 tmax_c = __ffma2_rn(make_float2(ltmax, rtmax), make_float2(slab_rel, slab_rel), make_float2(slab_abs, slab_abs));
			const float ltmax_c = tmax_c.x, rtmax_c = tmax_c.y;
#else
			const float ltmax_c = fmaf(ltmax, slab_rel, slab_abs), rtmax_c = fmaf(rtmax, slab_rel, slab_abs);
#endif
			const bool hl = ltmin <= ltmax_c && ltmin < tbest && ltmax_c > 0.0f;
			const bool hr = rtmin <= rtmax_c && rtmin < tbest && rtmax_c > 0.0f;
#else
			const bool hl = ltmin <= ltmax && ltmin < tbest && ltmax > 0.0f;
			const bool hr = rtmin <= rtmax && rtmin < tbest && rtmax > 0.0f;
#endif
			if (hl && hr) {
				// nearer child next, farther child on the stack (BVH.cu:91-97); boxes that both contain
				// the origin are ordered by where the ray leaves them
				const float lk = fmaxf(ltmin, 0.0f), rk = fmaxf(rtmin, 0.0f);
				const bool sw = lk > rk || (lk == rk && ltmax > rtmax);
				RTB_CHECK(sp >= 0 && sp < stack_cap, RTB_BOUNDS_STACK);
				STACK_PUSH(sw ? n3.x : n3.y);
				cur = sw ? n3.y : n3.x;
			} else if (hl) cur = n3.x;
			else if (hr) cur = n3.y;
			else if (sp > 0) { STACK_POP(cur); }
			else cur = TRAV_END;

#if !TRAV_WHILE_WHILE
			break;   // if-if flavour: at most one inner node per trip
#endif
		}
		if (cur == TRAV_END) break;
		if (cur < 0) {
			if (STATS) ++n_leaf;
			const int code = ~cur;
			int hit_code;
			const float t = leaf_test<(MEDIA == 2)>(sv, code, o, d, a, time, mr, tbest, RaySlab{idx, idy, idz, oix, oiy, oiz}, hit_code);
			if (t < tbest) { tbest = t; best = hit_code; }   // "if (t >= rec.distance) return false"  SphereHittable.cu:58
			if (sp == 0) break;
			STACK_POP(cur);
		}
	}
	if (STATS) { stats_out[0] = n_inner; stats_out[1] = n_leaf; }
	tbest_out = tbest; code_out = best;
#undef STACK_PUSH
#undef STACK_POP
}

#ifndef TRAVERSE_MIN_BLOCKS
#define TRAVERSE_MIN_BLOCKS 8   // 8 x 128 threads x 64 registers = the whole register file
#endif
// STACK = entries of the per-thread stack in shared memory: 16 when the tree is shallow enough (8 KB per
// block instead of 16 KB leaves 64 KB more L1 per SM: -2 % on the Book 2 final scene), 32 otherwise.
template <int MEDIA, int STACK>
__global__ void __launch_bounds__(TRAVERSE_THREADS, TRAVERSE_MIN_BLOCKS)
traverse_kernel(SceneView sv, BatchParams bp_in, WaveView wv, uint32_t bounce, int q) {
	const BatchParams bp = with_call_params(bp_in, wv);
	__shared__ int s_stack[STACK * TRAVERSE_THREADS];
	if (bounce >= *wv.tail_from) return;
	const uint32_t n = wv.n_live[bounce];
	if (n == 0) return;
	if (blockIdx.x == 0 && threadIdx.x == 0) { RTB_COUNT_CHECKS(n); RTB_CHECK(n <= wv.capacity, RTB_BOUNDS_QUEUE); }
	const uint32_t batch = bp.batch_base + *wv.batch_index * bp.batch_stride;
	const int lane = threadIdx.x & 31;
	const float4* __restrict__ rod = q ? wv.ray_od[1] : wv.ray_od[0];
	uint32_t* counter = wv.work + 2 * bounce;
	// Every warp owns one static chunk of 32 rays; only when the queue is longer than the whole
	// grid do warps pull further chunks from the device counter (short queues cost no atomics).
	const uint32_t warps_total = gridDim.x * (TRAVERSE_THREADS / 32);
	uint32_t base = (blockIdx.x * (TRAVERSE_THREADS / 32) + (threadIdx.x >> 5)) * 32u;
	const uint32_t static_span = warps_total * 32u;
	while (base < n) {
		const uint32_t i = base + lane;
		if (i < n) {
			float4 fo, fd; ld_ray(rod + 2 * (size_t)i, fo, fd);
			MediumRng mr = make_medium_rng(bp, batch, __float_as_uint(fd.w), bounce);
			float t; int code;
			trace_ray<MEDIA>(sv, xyz(fo), xyz(fd), fo.w, mr, s_stack + threadIdx.x, STACK, t, code);
			RTB_CHECK(i < wv.capacity, RTB_BOUNDS_QUEUE);
			stq(wv.hit + i, make_int2(__float_as_int(t), code));
		}
		__syncwarp();
		if (n <= static_span) break;
		uint32_t nb = 0;
		if (lane == 0) nb = atomicAdd(counter, 32u);
		base = static_span + __shfl_sync(FULL_MASK, nb, 0);
	}
}

// ------------------------------------------------------------------------------------------------
// surface reconstruction + textures + materials

struct Surface { v3 p, n_shade, n_geom; float u, v; };

// What Sphere::getNormal / MovingSphere::getNormal hand to materials: the outward normal
// (p - c) / r, never flipped (SphereHittable.cu:43-50,64,100).  Quads face the ray (book).
__device__ __forceinline__ void reconstruct(const SceneView& sv, int code, v3 o, v3 d, float time, float t, bool want_uv, Surface& s) {
	const int type = code & 15, base = type & 7;
	const bool xf = (type & PRIM_XF) != 0;
	RTB_CHECK(code >= 0 && (code >> RTB_LEAF_TYPE_BITS) + (xf && base >= PRIM_QUAD ? 1 : 0) < sv.n_prims, RTB_BOUNDS_PRIM);
	const float4* pp = sv.prims + 4 * (size_t)(code >> RTB_LEAF_TYPE_BITS);
	const float4 q0 = ldg4(pp);
	s.p = rt::madd(d, t, o);          // Material::Scatter uses in_ray.at(rec.distance): the world ray
	s.u = 0.0f; s.v = 0.0f;
	if (type == PRIM_MEDIUM_SPHERE || type == PRIM_MEDIUM_BOX) {   // normal arbitrary (book constant_medium::hit)
		s.n_geom = rt::mk(1.0f, 0.0f, 0.0f);
		s.n_shade = s.n_geom;
		return;
	}
	v3 oo = o, dd = d;                // the ray in the primitive's own frame
	const float4* tp = pp + xf_offset(base);
	if (xf) xf_ray(tp, o, d, oo, dd);
	if (base == PRIM_SPHERE || base == PRIM_MOVING_SPHERE) {
		v3 c = xyz(q0);
		if (base == PRIM_MOVING_SPHERE) c = rt::mix(c, xyz(ldg4(pp + 1)), time);
		v3 n = rt::divs(rt::sub(rt::madd(dd, t, oo), c), q0.w);   // (ray.at(t) - center) / radius   SphereHittable.cu:64,100
		if (want_uv) {   // book sphere::get_sphere_uv on the outward normal in the sphere's own frame
			float theta = acosf(-n.y);
			float phi = atan2f(-n.z, n.x) + 3.14159265358979323846f;
			s.u = phi / 6.28318530717958647692f;
			s.v = theta / 3.14159265358979323846f;
		}
		if (xf) n = xf_vec_to_world(tp, n);
		s.n_geom = n; s.n_shade = n;
	} else {
		const float4 q1 = ldg4(pp + 1), q2 = ldg4(pp + 2), q3 = ldg4(pp + 3);
		v3 N = rt::mk(q1.w, q2.w, q3.w);
		v3 ns = rt::dot(dd, N) > 0.0f ? rt::neg(N) : N;             // book set_face_normal
		if (want_uv) {
			v3 planar = rt::sub(rt::madd(dd, t, oo), xyz(q0));
			s.u = rt::dot(xyz(q3), rt::cross(planar, xyz(q2)));
			s.v = rt::dot(xyz(q3), rt::cross(xyz(q1), planar));
		}
		if (xf) { N = xf_vec_to_world(tp, N); ns = xf_vec_to_world(tp, ns); }
		s.n_geom = N; s.n_shade = ns;
	}
}

__device__ __forceinline__ float sin_any(float x) {
	float r = x * 0.15915494309189535f;
	r = r - floorf(r);
	float s, c; rt::sincos2pi(r, s, c);
	return s;
}

// Book perlin::noise with Hermite-smoothed trilinear interpolation of gradient dot products.
__device__ float perlin_noise(const float* __restrict__ grad, const int* __restrict__ perm, v3 p) {
	float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
	float u = p.x - fx, v = p.y - fy, w = p.z - fz;
	int i = (int)fx, j = (int)fy, k = (int)fz;
	float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
	float accum = 0.0f;
#pragma unroll
	for (int di = 0; di < 2; ++di)
#pragma unroll
		for (int dj = 0; dj < 2; ++dj)
#pragma unroll
			for (int dk = 0; dk < 2; ++dk) {
				int g = perm[(i + di) & 255] ^ perm[256 + ((j + dj) & 255)] ^ perm[512 + ((k + dk) & 255)];
				v3 c = rt::mk(grad[3 * g], grad[3 * g + 1], grad[3 * g + 2]);
				v3 wv = rt::mk(u - (float)di, v - (float)dj, w - (float)dk);
				float wi = di ? uu : 1.0f - uu, wj = dj ? vv : 1.0f - vv, wk = dk ? ww : 1.0f - ww;
				accum = fmaf(wi * wj * wk, rt::dot(c, wv), accum);
			}
	return accum;
}

__device__ v3 texture_value(const SceneView& sv, int tex, float u, float v, v3 p) {
	for (int guard = 0; guard < 16; ++guard) {
		RTB_CHECK(tex >= 0 && tex < sv.n_textures, RTB_BOUNDS_TEXTURE);
		const float4 t0 = ldg4(sv.textures + 3 * tex);
		const int kind = __float_as_int(t0.x);
		if (kind == RTB_TEX_SOLID) { return xyz(ldg4(sv.textures + 3 * tex + 1)); }
		if (kind == RTB_TEX_CHECKER) {
			// checker_texture::value  cu_Textures.cuh:32-39: C cast (truncation) and C '%'
			float inv = t0.w;
			int sum = (int)(p.x * inv) + (int)(p.y * inv) + (int)(p.z * inv);
			tex = (sum % 2 == 0) ? __float_as_int(t0.y) : __float_as_int(t0.z);
			continue;
		}
		const float4 t2 = ldg4(sv.textures + 3 * tex + 2);
		const int w = __float_as_int(t2.x), h = __float_as_int(t2.y);
		const uint32_t off = __float_as_uint(t2.z);
		if (kind == RTB_TEX_IMAGE) {   // book image_texture::value (nearest texel, bytes / 255)
			float uc = fminf(fmaxf(u, 0.0f), 1.0f), vc = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
			int i = (int)(uc * (float)w), j = (int)(vc * (float)h);
			i = i < w - 1 ? i : w - 1; j = j < h - 1 ? j : h - 1;
			RTB_CHECK(i >= 0 && j >= 0 && off + 3 * ((size_t)j * w + i) + 2 < sv.n_blob, RTB_BOUNDS_TEXTURE);
			const uint8_t* px = sv.blob + off + 3 * ((size_t)j * w + i);
			const float sc = 1.0f / 255.0f;
			return rt::mk(sc * (float)px[0], sc * (float)px[1], sc * (float)px[2]);
		}
		// RTB_TEX_NOISE: book noise_texture::value = 0.5 * (1 + sin(scale * p.z + 10 * turb(p, 7)))
		const float* grad = reinterpret_cast<const float*>(sv.blob + off);
		const int* perm = reinterpret_cast<const int*>(sv.blob + off + 256 * 3 * 4);
		float accum = 0.0f, weight = 1.0f; v3 tp = p;
		for (int k = 0; k < 7; ++k) {
			accum = fmaf(weight, perlin_noise(grad, perm, tp), accum);
			weight *= 0.5f; tp = rt::mul(tp, 2.0f);
		}
		float turb = fabsf(accum);
		float val = 0.5f * (1.0f + sin_any(fmaf(t0.w, p.z, 10.0f * turb)));
		return rt::mk(val, val, val);
	}
	return rt::mk(0.0f, 0.0f, 0.0f);
}

__device__ __forceinline__ bool texture_needs_uv(const SceneView& sv, int tex) {
	if (tex < 0) return false;
	const int kind = __float_as_int(ldg4(sv.textures + 3 * tex).x);
	return kind != RTB_TEX_SOLID && kind != RTB_TEX_NOISE;   // checker children may be images
}

// Material::Scatter for every material kind.  Returns 0 = absorbed (path ends black),
// 1 = scattered (dir/attenuation valid), 2 = emitted (attenuation holds the emitted radiance).
// With DEFER, an image / noise albedo of a scattering material is not evaluated here: attenuation is
// left at 1 and defer_tex names the texture, to be multiplied in later by texture_kernel (dense warps).
template <bool DEFER>
__device__ __forceinline__ int scatter(const SceneView& sv, int mat, const Surface& s, v3 d, const rt::f4& r, v3& dir, v3& att, int& defer_tex) {
	RTB_CHECK(mat >= 0 && mat < sv.n_materials, RTB_BOUNDS_MATERIAL);
	const float4 m0 = ldg4(sv.materials + 2 * mat), m1 = ldg4(sv.materials + 2 * mat + 1);
	const int kind = __float_as_int(m0.x), tex = __float_as_int(m0.y);
	const float param = m0.z;
	defer_tex = -1;
	v3 albedo;
	if (tex < 0) albedo = xyz(m1);
	else {
		const int tkind = __float_as_int(ldg4(sv.textures + 3 * tex).x);
		if (DEFER && kind != RTB_MAT_DIFFUSE_LIGHT && (tkind == RTB_TEX_NOISE || tkind == RTB_TEX_IMAGE)) { defer_tex = tex; albedo = rt::mk(1.0f, 1.0f, 1.0f); }
		else albedo = texture_value(sv, tex, s.u, s.v, s.p);
	}
	switch (kind) {
	case RTB_MAT_LAMBERTIAN: {   // LambertianAbstract / LambertianTexture  cu_materials.cuh:26-40,52-64
		dir = rt::add(s.n_shade, rt::unit_sphere(r.x, r.y));
		if (rt::near_zero(dir, 1e-9f)) return 0;
		att = albedo; return 1;
	}
	case RTB_MAT_METAL: {        // MetalAbstract  cu_materials.cuh:77-95 (in_ray.d is not normalised)
		dir = rt::madd(rt::unit_sphere(r.x, r.y), param, rt::reflect(d, s.n_shade));
		if (rt::dot(dir, s.n_shade) < 0.0f || rt::near_zero(dir, 1e-9f)) return 0;
		att = albedo; return 1;
	}
	case RTB_MAT_DIELECTRIC: {   // DielectricAbstract  cu_materials.cuh:115-143, reflectance :99-104
		v3 n = s.n_geom;
		bool back = rt::dot(d, n) > 0.0f;          // isBackfacing  ray_data.cuh:44-46
		if (back) n = rt::neg(n);
		float ratio = back ? param : 1.0f / param;
		v3 ud = rt::normalize(d);
		float cos_theta = fminf(rt::dot(rt::neg(ud), n), 1.0f);
		float sin_theta = sqrtf(fmaf(-cos_theta, cos_theta, 1.0f));
		float r0 = (1.0f - ratio) / (1.0f + ratio); r0 = r0 * r0;
		float x = 1.0f - cos_theta, x2 = x * x;
		float prob = fmaf(1.0f - r0, x2 * x2 * x, r0);   // powf(1 - cos, 5) as exact products
		if (ratio * sin_theta > 1.0f || prob > r.z) {
			dir = rt::reflect(ud, n);
		} else {                                   // glm::refract  func_geometric.inl:113-124
			float dv = rt::dot(n, ud);
			float k = fmaf(-(ratio * ratio), fmaf(-dv, dv, 1.0f), 1.0f);
			if (k < 0.0f) return 0;
			float coef = fmaf(ratio, dv, sqrtf(k));
			dir = rt::mk(fmaf(-coef, n.x, ratio * ud.x), fmaf(-coef, n.y, ratio * ud.y), fmaf(-coef, n.z, ratio * ud.z));
		}
		if (dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f) return 0;
		att = albedo; return 1;
	}
	case RTB_MAT_ISOTROPIC: {    // book isotropic::scatter
		dir = rt::unit_sphere(r.x, r.y);
		att = albedo; return 1;
	}
	default:                     // RTB_MAT_DIFFUSE_LIGHT: emits, never scatters (book diffuse_light)
		att = albedo; return 2;
	}
}

__device__ __forceinline__ v3 background(const SceneView& sv, v3 d) {
	if (sv.background_mode == RTB_BG_CONSTANT) return rt::mk(sv.bg_r, sv.bg_g, sv.bg_b);
	// Renderer.cu:149-151: t = normalize(d).y * 0.5 + 0.5; lerp((0.1,0.2,0.4), (0.9,0.9,0.99), t)
	float ny = d.y * (1.0f / sqrtf(rt::dot(d, d)));
	float t = fmaf(ny, 0.5f, 0.5f);
	return rt::mk(fmaf(0.9f - 0.1f, t, 0.1f), fmaf(0.9f - 0.2f, t, 0.2f), fmaf(0.99f - 0.4f, t, 0.4f));
}

// ------------------------------------------------------------------------------------------------
// shade: one path segment of sample_world (Renderer.cu:146-176) + block-level queue compaction

// Shades one traversed segment.  Returns true when the path continues (no/nd/nthr hold the scattered
// ray and the updated throughput); otherwise the path has ended and its contribution (if any) has
// been written to contrib[path].
struct DeferredTex { int tex; float u, v; v3 p; };

template <bool DEFER>
__device__ __forceinline__ bool shade_segment(const SceneView& sv, const BatchParams& bp, const WaveView& wv, uint32_t batch, uint32_t bounce,
                                              const float4& fo, const float4& fd, v3 thr, float t, int code,
                                              float4& no, float4& nd, v3& nthr, DeferredTex& dt) {
	dt.tex = -1;
	const uint32_t path = __float_as_uint(fd.w);
	const v3 o = xyz(fo), d = xyz(fd);
	if (code < 0) {
		v3 c = rt::mulv(thr, background(sv, d));          // miss: throughput * sky   Renderer.cu:153
		wv.contrib[path] = make_float4(c.x, c.y, c.z, 0.0f);
		return false;
	}
	RTB_CHECK((code >> RTB_LEAF_TYPE_BITS) < sv.n_prims, RTB_BOUNDS_PRIM);
	const int2 info = __ldg(sv.prim_info + (code >> RTB_LEAF_TYPE_BITS));
	RTB_CHECK(info.x >= 0 && info.x < sv.n_materials, RTB_BOUNDS_MATERIAL);
	RTB_CHECK(path < wv.capacity, RTB_BOUNDS_PATH);
	const int tex = __float_as_int(ldg4(sv.materials + 2 * info.x).y);
	Surface s;
	reconstruct(sv, code, o, d, fo.w, t, texture_needs_uv(sv, tex), s);
	uint32_t pixel, sample;
	path_pixel_sample(bp, batch, path, pixel, sample);
	const rt::f4 r = rt::rng4(bp.seed, pixel, sample, bounce, rt::STREAM_SCATTER);
	v3 dir, att;
	const int res = scatter<DEFER>(sv, info.x, s, d, r, dir, att, dt.tex);
	dt.u = s.u; dt.v = s.v; dt.p = s.p;
	if (res == 2) {                                       // emitter: throughput * emitted, path ends
		v3 c = rt::mulv(thr, att);
		wv.contrib[path] = make_float4(c.x, c.y, c.z, 0.0f);
		return false;
	}
	if (res != 1 || bounce + 1 >= bp.max_depth) { dt.tex = -1; return false; }   // absorbed, or out of depth: black   Renderer.cu:164,180
	// scatter_ray = Ray(at(t), dir, time); o += d * 0.001   Renderer.cu:168-175
	v3 o2 = rt::madd(dir, 0.001f, s.p);
	nthr = rt::mulv(thr, att);
	no = make_float4(o2.x, o2.y, o2.z, fo.w);
	nd = make_float4(dir.x, dir.y, dir.z, fd.w);
	return true;
}

// Bin of a ray for the binning pass: Morton code of the origin's cell on a 2^org_bits cube over the world bounds, then the
// cell of the direction on a 2^dir_bits square of the octahedral map.  Rays of one bin start close together and head the
// same way, so the lanes of a warp walk the same nodes (measured on B200: 12 -> 20+ active lanes on secondary bounces).
__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // 10 bits -> every third bit
	v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
	return v;
}
__device__ __forceinline__ uint32_t ray_bin(const SceneView& sv, const float4& o, const float4& d) {
	const int qo = 1 << sv.bin_org_bits;
	int ix = (int)((o.x - sv.bin_min[0]) * sv.bin_scale[0]), iy = (int)((o.y - sv.bin_min[1]) * sv.bin_scale[1]), iz = (int)((o.z - sv.bin_min[2]) * sv.bin_scale[2]);
	ix = min(max(ix, 0), qo - 1); iy = min(max(iy, 0), qo - 1); iz = min(max(iz, 0), qo - 1);
	uint32_t key = spread3((uint32_t)ix) | (spread3((uint32_t)iy) << 1) | (spread3((uint32_t)iz) << 2);
	if (sv.bin_dir_bits > 0) {
		const float s = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
		float u = d.x * s, v = d.z * s;
		if (d.y < 0.0f) { const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f); u = uu; v = vv; }
		const int qd = 1 << sv.bin_dir_bits;
		int iu = (int)((u * 0.5f + 0.5f) * (float)qd), iv = (int)((v * 0.5f + 0.5f) * (float)qd);
		iu = min(max(iu, 0), qd - 1); iv = min(max(iv, 0), qd - 1);
		key = (key << (2 * sv.bin_dir_bits)) | (uint32_t)(iv * qd + iu);
	}
	return key;
}
// Shared-memory hash table of (bin, rays) used by the binning kernels (see "binning" below).
// (B200, r2: tiles of 1,024 / 2,048 / 4,096 rays = 256 / 512 / 1,024 threads x 4: binning 4.75 / 4.05 / 4.17 ms per batch - a global
// counter is touched once per distinct key of a TILE, so larger tiles send fewer atomics, until one block per SM is too few.)
#ifndef BIN_THREADS
#define BIN_THREADS 512
#endif
#ifndef BIN_ITEMS
#define BIN_ITEMS 4                                  // rays per thread per tile
#endif
#define BIN_TILE (BIN_THREADS * BIN_ITEMS)
#ifndef BIN_SLOTS_LOG2
#define BIN_SLOTS_LOG2 12
#endif
#ifndef BIN_MIN_BLOCKS
#define BIN_MIN_BLOCKS 2                             // 2 x 512 threads x 64 registers
#endif
#define BIN_LAUNCH_BOUNDS __launch_bounds__(BIN_THREADS, BIN_MIN_BLOCKS)
#define BIN_SLOTS (1 << BIN_SLOTS_LOG2)              // hash slots per block: at most half full with one tile's keys
#define BIN_EMPTY 0xFFFFFFFFu
#define BIN_COUNT_SMEM (2 * BIN_SLOTS * sizeof(uint32_t))
#define BIN_PERMUTE_SMEM (3 * BIN_SLOTS * sizeof(uint32_t))

// Finds or claims the slot of `key` and counts one ray in it (linear probing; callers keep the table at most half full).
// Returns the slot, with the top bit set when this call claimed it.
__device__ __forceinline__ uint32_t bin_table_add(uint32_t* s_key, uint32_t* s_cnt, uint32_t key) {
	uint32_t h = (key * 2654435761u) >> (32 - BIN_SLOTS_LOG2);  // Fibonacci hash to log2(BIN_SLOTS) bits
	for (;;) {
		const uint32_t prev = atomicCAS(s_key + h, BIN_EMPTY, key);
		if (prev == BIN_EMPTY || prev == key) { atomicAdd(s_cnt + h, 1u); return prev == BIN_EMPTY ? (h | 0x80000000u) : h; }
		h = (h + 1u) & (BIN_SLOTS - 1u);
	}
}

// One global atomic per bin the table holds, then the table is empty again (all threads of the block).
__device__ __forceinline__ void bin_table_flush(uint32_t* s_key, uint32_t* s_cnt, uint32_t* bin_count, uint32_t n_bins, int threads) {
	for (uint32_t s = threadIdx.x; s < BIN_SLOTS; s += threads) {
		const uint32_t key = s_key[s];
		if (key != BIN_EMPTY) { RTB_CHECK(key < n_bins, RTB_BOUNDS_BIN); atomicAdd(bin_count + key, s_cnt[s]); s_key[s] = BIN_EMPTY; s_cnt[s] = 0u; }
	}
}

#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 8   // 8 x 128 threads x 64 registers = the whole register file (6 / 9 / 10 blocks: 11.7 / 10.9 / 12.4 ms)
#endif
__global__ void __launch_bounds__(SHADE_THREADS, SHADE_MIN_BLOCKS)
shade_kernel(SceneView sv, BatchParams bp_in, WaveView wv, uint32_t bounce, int q) {
	const BatchParams bp = with_call_params(bp_in, wv);
	__shared__ uint32_t s_chunk, s_base;
	__shared__ uint32_t s_warp[SHADE_THREADS / 32];
	if (bounce >= *wv.tail_from) return;
	const uint32_t n = wv.n_live[bounce];
	if (n == 0) return;
	const uint32_t batch = bp.batch_base + *wv.batch_index * bp.batch_stride;
	const int in = q, out = in ^ 1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const float4* __restrict__ rod = in ? wv.ray_od[1] : wv.ray_od[0];
	const float4* __restrict__ rt_ = in ? wv.thr[1] : wv.thr[0];
	float4* __restrict__ wod = out ? wv.ray_od[1] : wv.ray_od[0];
	float4* __restrict__ wt = out ? wv.thr[1] : wv.thr[0];
	uint32_t* counter = wv.work + 2 * bounce + 1;

	// the first chunk of every block is static; further chunks come from the device counter
	uint32_t base = blockIdx.x * SHADE_THREADS;
	const uint32_t static_span = gridDim.x * SHADE_THREADS;
	for (;;) {
		if (base >= n) break;
		const uint32_t i = base + threadIdx.x;
		bool alive = false;
		float4 no = make_float4(0, 0, 0, 0), nd = no; v3 nthr = rt::mk(0, 0, 0);
		DeferredTex dt; dt.tex = -1; dt.u = dt.v = 0.0f; dt.p = rt::mk(0, 0, 0);
		if (i < n) {
			float4 fo, fd; ld_ray(rod + 2 * (size_t)i, fo, fd);
			const float4 ft = ldq(rt_ + i);
			const int2 h = ldq(wv.hit + i);
			alive = shade_segment<true>(sv, bp, wv, batch, bounce, fo, fd, xyz(ft), __int_as_float(h.x), h.y, no, nd, nthr, dt);
		}
		// live-path compaction: warp ballot/popc, one global atomic per block
		const uint32_t mask = __ballot_sync(FULL_MASK, alive);
		if (lane == 0) s_warp[warp] = __popc(mask);
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t tot = 0;
#pragma unroll
			for (int w = 0; w < SHADE_THREADS / 32; ++w) { uint32_t c = s_warp[w]; s_warp[w] = tot; tot += c; }
			s_base = tot ? atomicAdd(wv.n_live + bounce + 1, tot) : 0u;
			s_chunk = (n > static_span) ? static_span + atomicAdd(counter, (uint32_t)SHADE_THREADS) : n;
		}
		__syncthreads();
		if (alive) {
			const uint32_t pos = s_base + s_warp[warp] + __popc(mask & ((1u << lane) - 1u));
			RTB_CHECK(pos < wv.capacity && pos < n, RTB_BOUNDS_QUEUE);
			st_ray(wod + 2 * (size_t)pos, no, nd); stq(wt + pos, make_float4(nthr.x, nthr.y, nthr.z, 0.0f));
			// texture work list: (p, queue slot), (u, v, -, texture id)
			const bool defer = dt.tex >= 0;
			const uint32_t act = __activemask();
			const uint32_t dmask = __ballot_sync(act, defer);
			if (dmask) {
				const int leader = __ffs(dmask) - 1;
				uint32_t tb = 0;
				if (lane == leader) tb = atomicAdd(wv.n_tex + bounce, (uint32_t)__popc(dmask));
				tb = __shfl_sync(act, tb, leader);
				if (defer) {
					const uint32_t k = tb + __popc(dmask & ((1u << lane) - 1u));
					wv.tex_work[2 * (size_t)k] = make_float4(dt.p.x, dt.p.y, dt.p.z, __uint_as_float(pos));
					wv.tex_work[2 * (size_t)k + 1] = make_float4(dt.u, dt.v, 0.0f, __int_as_float(dt.tex));
				}
			}
		}
		base = s_chunk;
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// texture: evaluates the image / Perlin albedos that shade deferred, with dense warps, and folds them
// into the throughput of the scattered rays (thr * value is the same single product either way).

__global__ void __launch_bounds__(STREAM_THREADS)
texture_kernel(SceneView sv, WaveView wv, uint32_t bounce, int q_out) {
	if (bounce >= *wv.tail_from) return;
	const uint32_t n = wv.n_tex[bounce];
	if (n == 0) return;
	float4* __restrict__ wt = q_out ? wv.thr[1] : wv.thr[0];
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const float4 a = wv.tex_work[2 * (size_t)k], b = wv.tex_work[2 * (size_t)k + 1];
		const v3 val = texture_value(sv, __float_as_int(b.w), b.x, b.y, xyz(a));
		const uint32_t pos = __float_as_uint(a.w);
		RTB_CHECK(pos < wv.capacity && __float_as_int(b.w) >= 0 && __float_as_int(b.w) < sv.n_textures, RTB_BOUNDS_QUEUE);
		float4 t = wt[pos];
		t.x *= val.x; t.y *= val.y; t.z *= val.z;
		wt[pos] = t;
	}
}

// ------------------------------------------------------------------------------------------------
// binning: a counting sort of the live queue by ray_bin between shade(b - 1) and traverse(b), in three launches.
//   bin_count    rays per bin
//   bin_scan     exclusive prefix sums of the counts = the first slot of every bin
//   bin_permute  every ray moves to the next free slot of its bin in the other queue
// Rays pile up in few bins (a third of the late-bounce rays of the Book 2 final scene scatter inside one sphere of
// smoke), so neither kernel sends one global atomic per ray: each block aggregates the keys of its rays in a small
// shared-memory hash table first - same-address atomics are cheap there - and touches a global counter once per distinct
// key.  Where a ray sits in a queue changes nothing about the image - contributions are stored by path id and summed per
// pixel in sample order - only which rays share a warp.

// Rays per bin.  A block's counts gather in its table over many tiles and go to the global counters when the table might
// not hold another tile's keys, and at the end.
__global__ void BIN_LAUNCH_BOUNDS
bin_count_kernel(SceneView sv, WaveView wv, uint32_t bounce, int q) {
	extern __shared__ uint32_t s_bin[];                   // BIN_COUNT_SMEM bytes
	uint32_t* const s_key = s_bin; uint32_t* const s_cnt = s_bin + BIN_SLOTS;
	__shared__ uint32_t s_used;
	if (bounce >= *wv.tail_from) return;
	const uint32_t n = wv.n_live[bounce];
	if (n == 0) return;
	const float4* __restrict__ rod = q ? wv.ray_od[1] : wv.ray_od[0];
	for (uint32_t s = threadIdx.x; s < BIN_SLOTS; s += BIN_THREADS) { s_key[s] = BIN_EMPTY; s_cnt[s] = 0u; }
	if (threadIdx.x == 0) s_used = 0u;
	__syncthreads();
	for (uint32_t base = blockIdx.x * BIN_TILE; base < n; base += gridDim.x * BIN_TILE) {
		uint32_t claimed = 0;
#pragma unroll
		for (int k = 0; k < BIN_ITEMS; ++k) {
			const uint32_t i = base + k * BIN_THREADS + threadIdx.x;
			if (i < n) { float4 o, d; ld_ray(rod + 2 * (size_t)i, o, d); claimed += bin_table_add(s_key, s_cnt, ray_bin(sv, o, d)) >> 31; }
		}
		if (claimed) atomicAdd(&s_used, claimed);
		__syncthreads();
		if (s_used > BIN_SLOTS / 2 - BIN_TILE / 2) {          // (a tile adds at most BIN_TILE keys; past this mark, flush)
			__syncthreads();
			bin_table_flush(s_key, s_cnt, wv.bin_count, wv.n_bins, BIN_THREADS);
			if (threadIdx.x == 0) s_used = 0u;
			__syncthreads();
		}
	}
	bin_table_flush(s_key, s_cnt, wv.bin_count, wv.n_bins, BIN_THREADS);
}

// Block b owns bins [b * 4096, (b + 1) * 4096): it adds up everything before them (the counts are re-read from L2 - at
// most 1 MB - instead of passing block totals around), then scans its own.  The counters are zeroed by bin_permute.
#define BIN_SCAN_THREADS 1024
#define BIN_SCAN_PER_BLOCK 4096
__global__ void __launch_bounds__(BIN_SCAN_THREADS)
bin_scan_kernel(WaveView wv, uint32_t bounce) {
	__shared__ uint32_t s_part[BIN_SCAN_THREADS];
	__shared__ uint32_t s_before;
	if (bounce >= *wv.tail_from || wv.n_live[bounce] == 0) return;
	const uint32_t first = blockIdx.x * BIN_SCAN_PER_BLOCK;
	uint32_t sum = 0;
	for (uint32_t b = threadIdx.x; b < first; b += BIN_SCAN_THREADS) sum += wv.bin_count[b];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	for (uint32_t off = BIN_SCAN_THREADS / 2; off > 0; off >>= 1) {
		if (threadIdx.x < off) s_part[threadIdx.x] += s_part[threadIdx.x + off];
		__syncthreads();
	}
	if (threadIdx.x == 0) s_before = s_part[0];
	__syncthreads();
	// own bins: 4 consecutive bins per thread (one 16-byte load), block-wide inclusive scan of the per-thread sums
	const uint32_t b0 = first + 4u * threadIdx.x;
	uint4 c = make_uint4(0, 0, 0, 0);
	if (b0 + 3u < wv.n_bins) c = *reinterpret_cast<const uint4*>(wv.bin_count + b0);
	else { if (b0 < wv.n_bins) c.x = wv.bin_count[b0]; if (b0 + 1u < wv.n_bins) c.y = wv.bin_count[b0 + 1u]; if (b0 + 2u < wv.n_bins) c.z = wv.bin_count[b0 + 2u]; }
	const uint32_t mine = c.x + c.y + c.z + c.w;
	__syncthreads();
	s_part[threadIdx.x] = mine;
	__syncthreads();
	for (uint32_t off = 1; off < BIN_SCAN_THREADS; off <<= 1) {
		const uint32_t v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
		__syncthreads();
		s_part[threadIdx.x] += v;
		__syncthreads();
	}
	uint32_t run = s_before + s_part[threadIdx.x] - mine;
	if (b0 < wv.n_bins) wv.bin_cursor[b0] = run; run += c.x;
	if (b0 + 1u < wv.n_bins) wv.bin_cursor[b0 + 1u] = run; run += c.y;
	if (b0 + 2u < wv.n_bins) wv.bin_cursor[b0 + 2u] = run; run += c.z;
	if (b0 + 3u < wv.n_bins) wv.bin_cursor[b0 + 3u] = run;
}

__global__ void BIN_LAUNCH_BOUNDS
bin_permute_kernel(SceneView sv, WaveView wv, uint32_t bounce, int q_from) {
	extern __shared__ uint32_t s_bin[];                   // BIN_PERMUTE_SMEM bytes
	uint32_t* const s_key = s_bin; uint32_t* const s_cnt = s_bin + BIN_SLOTS; uint32_t* const s_base = s_bin + 2 * BIN_SLOTS;
	if (bounce >= *wv.tail_from) return;
	const uint32_t n = wv.n_live[bounce];
	if (n == 0) return;
	// the counters of this bounce are dead once bin_scan has run: leave them zero for the next binned bounce
	for (uint32_t b = blockIdx.x * BIN_THREADS + threadIdx.x; b < wv.n_bins; b += gridDim.x * BIN_THREADS) wv.bin_count[b] = 0u;
	const float4* __restrict__ rod = q_from ? wv.ray_od[1] : wv.ray_od[0];
	const float4* __restrict__ rt_ = q_from ? wv.thr[1] : wv.thr[0];
	float4* __restrict__ wod = q_from ? wv.ray_od[0] : wv.ray_od[1];
	float4* __restrict__ wt = q_from ? wv.thr[0] : wv.thr[1];
	for (uint32_t s = threadIdx.x; s < BIN_SLOTS; s += BIN_THREADS) { s_key[s] = BIN_EMPTY; s_cnt[s] = 0u; }
	__syncthreads();
	for (uint32_t base = blockIdx.x * BIN_TILE; base < n; base += gridDim.x * BIN_TILE) {
		float4 o[BIN_ITEMS], d[BIN_ITEMS], t[BIN_ITEMS];
		uint32_t slot[BIN_ITEMS];
#pragma unroll
		for (int k = 0; k < BIN_ITEMS; ++k) {
			const uint32_t i = base + k * BIN_THREADS + threadIdx.x;
			if (i < n) { ld_ray(rod + 2 * (size_t)i, o[k], d[k]); t[k] = ldq(rt_ + i); }
		}
#pragma unroll
		for (int k = 0; k < BIN_ITEMS; ++k) {
			const uint32_t i = base + k * BIN_THREADS + threadIdx.x;
			slot[k] = i < n ? (bin_table_add(s_key, s_cnt, ray_bin(sv, o[k], d[k])) & 0x7FFFFFFFu) : 0u;
		}
		__syncthreads();
		// a run of slots in the other queue for every distinct key of the tile
		for (uint32_t s = threadIdx.x; s < BIN_SLOTS; s += BIN_THREADS) {
			const uint32_t key = s_key[s];
			if (key != BIN_EMPTY) { RTB_CHECK(key < wv.n_bins, RTB_BOUNDS_BIN); s_base[s] = atomicAdd(wv.bin_cursor + key, s_cnt[s]); s_cnt[s] = 0u; }
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < BIN_ITEMS; ++k) {
			const uint32_t i = base + k * BIN_THREADS + threadIdx.x;
			if (i < n) {
				const uint32_t pos = s_base[slot[k]] + atomicAdd(s_cnt + slot[k], 1u);
				RTB_CHECK(pos < n && pos < wv.capacity, RTB_BOUNDS_QUEUE);
				st_ray(wod + 2 * (size_t)pos, o[k], d[k]); wt[pos] = t[k];   // (the 32-byte ray record is one whole sector)
			}
		}
		__syncthreads();
		for (uint32_t s = threadIdx.x; s < BIN_SLOTS; s += BIN_THREADS) { s_key[s] = BIN_EMPTY; s_cnt[s] = 0u; }
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// tail: once the live queue is too short to fill the machine, per-bounce launches are pure latency.
// One persistent launch then runs every remaining path to completion (traverse + shade fused, one
// thread per path); later traverse/shade launches of the batch see tail_from and return at once.

template <bool MEDIA, int STACK>
__global__ void __launch_bounds__(TRAVERSE_THREADS)
tail_kernel(SceneView sv, BatchParams bp_in, WaveView wv, uint32_t bounce0, int q, uint32_t threshold) {
	const BatchParams bp = with_call_params(bp_in, wv);
	__shared__ int s_stack[STACK * TRAVERSE_THREADS];
	if (*wv.tail_from < bounce0) return;                 // an earlier checkpoint already took the batch over
	const uint32_t n = wv.n_live[bounce0];
	if (n == 0 || n > threshold) return;
	if (blockIdx.x == 0 && threadIdx.x == 0) *wv.tail_from = bounce0;
	const uint32_t batch = bp.batch_base + *wv.batch_index * bp.batch_stride;
	const int in = q;
	const float4* __restrict__ rod = in ? wv.ray_od[1] : wv.ray_od[0];
	const float4* __restrict__ rt_ = in ? wv.thr[1] : wv.thr[0];
	unsigned long long extra = 0;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		float4 fo, fd; ld_ray(rod + 2 * (size_t)i, fo, fd);
		v3 thr = xyz(rt_[i]);
		MediumRng mr = make_medium_rng(bp, batch, __float_as_uint(fd.w), bounce0);
		for (uint32_t b = bounce0; b < bp.max_depth; ++b) {
			if (b > bounce0) ++extra;
			mr.bounce = b; mr.block = 0xFFFFFFFFu;
			float t; int code;
			trace_ray<(MEDIA ? 2 : 0)>(sv, xyz(fo), xyz(fd), fo.w, mr, s_stack + threadIdx.x, STACK, t, code);
			float4 no, nd; v3 nthr;
			DeferredTex dt;
			if (!shade_segment<false>(sv, bp, wv, batch, b, fo, fd, thr, t, code, no, nd, nthr, dt)) break;
			fo = no; fd = nd; thr = nthr;
		}
	}
	// ray segments traced beyond the first one of each path (that one is counted through n_live)
	for (int off = 16; off > 0; off >>= 1) extra += __shfl_down_sync(FULL_MASK, extra, off);
	if ((threadIdx.x & 31) == 0 && extra) atomicAdd(wv.totals + 1, extra);
}

// ------------------------------------------------------------------------------------------------
// accumulate + batch epilogue + resolve

__global__ void __launch_bounds__(STREAM_THREADS)
accumulate_kernel(BatchParams bp_in, WaveView wv, float4* __restrict__ accum, float4* __restrict__ accum2) {
	const BatchParams bp = with_call_params(bp_in, wv);
	const uint32_t batch = bp.batch_base + *wv.batch_index * bp.batch_stride;
	const uint32_t ns = batch_sample_count(bp, batch);
	if (ns == 0) return;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t pl = blockIdx.x * blockDim.x + threadIdx.x; pl < bp.npix; pl += stride) {
		float sx = 0.0f, sy = 0.0f, sz = 0.0f, qx = 0.0f, qy = 0.0f, qz = 0.0f;
		for (uint32_t s = 0; s < ns; ++s) {   // fixed sample order: deterministic sums
			const float4 c = wv.contrib[(size_t)s * bp.npix + pl];
			sx += c.x; sy += c.y; sz += c.z;
			if (bp.variance) { qx = fmaf(c.x, c.x, qx); qy = fmaf(c.y, c.y, qy); qz = fmaf(c.z, c.z, qz); }
		}
		const uint32_t gid = bp.row_begin * bp.width + pl;
		float4 a = accum[gid];
		a.x += sx; a.y += sy; a.z += sz; a.w += (float)ns;
		accum[gid] = a;
		if (bp.variance) {
			float4 q = accum2[gid];
			q.x += qx; q.y += qy; q.z += qz; q.w += (float)ns;
			accum2[gid] = q;
		}
	}
}

__global__ void end_batch_kernel(BatchParams bp_in, WaveView wv) {
	const BatchParams bp = with_call_params(bp_in, wv);
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	unsigned long long rays = 0;
	for (uint32_t b = 0; b < bp.max_depth; ++b) rays += wv.n_live[b];
	wv.totals[0] += wv.n_live[0];
	wv.totals[1] += rays;
	*wv.batch_index += 1;
}

__global__ void __launch_bounds__(STREAM_THREADS)
resolve_kernel(const float4* __restrict__ accum, float4* __restrict__ out, uint32_t n) {
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float4 a = accum[i];
		const float inv = 1.0f / a.w;                              // radiance *= 1.0f / spp   Renderer.cu:206
		float r = a.x * inv, g = a.y * inv, b = a.z * inv;
		r = fminf(fmaxf(r, 0.0f), 1.0f); g = fminf(fmaxf(g, 0.0f), 1.0f); b = fminf(fmaxf(b, 0.0f), 1.0f);
		out[i] = make_float4(sqrtf(r), sqrtf(g), sqrtf(b), 1.0f);  // clamp, gamma 2, alpha 1   :209-214
	}
}

// Output stage of FirstApp::write_renderbuffer (FirstApp.cpp:108-122) on the device: mean -> clamp -> sqrt (as resolve),
// uint8 = value * 255.999f, RGB, optionally with the rows flipped (the float buffer's row 0 is the bottom of the picture).
__global__ void __launch_bounds__(STREAM_THREADS)
quantize_kernel(const float4* __restrict__ accum, uint8_t* __restrict__ rgb, uint32_t width, uint32_t height, int flip_rows) {
	const uint32_t n = width * height, stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint32_t y = i / width, x = i - y * width;
		const float4 a = accum[(size_t)(flip_rows ? height - 1 - y : y) * width + x];
		const float inv = 1.0f / a.w;
		const float r = sqrtf(fminf(fmaxf(a.x * inv, 0.0f), 1.0f)), g = sqrtf(fminf(fmaxf(a.y * inv, 0.0f), 1.0f)), b = sqrtf(fminf(fmaxf(a.z * inv, 0.0f), 1.0f));
		rgb[3 * (size_t)i] = (uint8_t)(r * 255.999f); rgb[3 * (size_t)i + 1] = (uint8_t)(g * 255.999f); rgb[3 * (size_t)i + 2] = (uint8_t)(b * 255.999f);
	}
}

// ------------------------------------------------------------------------------------------------
// hit-record parity hook

template <int STACK>
__global__ void __launch_bounds__(TRAVERSE_THREADS)
trace_rays_kernel(SceneView sv, const float4* __restrict__ ro, const float4* __restrict__ rd, uint32_t n,
                  int2* __restrict__ hit, int2* __restrict__ stats, uint32_t* counter) {
	__shared__ int s_stack[STACK * TRAVERSE_THREADS];
	const int lane = threadIdx.x & 31;
	for (;;) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(counter, 32u);
		base = __shfl_sync(FULL_MASK, base, 0);
		if (base >= n) break;
		uint32_t i = base + lane;
		if (i < n) {
			float4 fo = ro[i], fd = rd[i];
			MediumRng mr = make_medium_rng(BatchParams{}, 0, 0, 0);
			float t; int code;
			int st[2];
			trace_ray<0, true>(sv, xyz(fo), xyz(fd), fo.w, mr, s_stack + threadIdx.x, STACK, t, code, st);
			hit[i] = make_int2(__float_as_int(t), code);
			stats[i] = make_int2(st[0], st[1]);
		}
		__syncwarp();
	}
}

__global__ void __launch_bounds__(STREAM_THREADS)
hit_record_kernel(SceneView sv, const float4* __restrict__ ro, const float4* __restrict__ rd, const int2* __restrict__ hit,
                  const int2* __restrict__ stats, uint32_t n, rtb_hit* __restrict__ out) {
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		rtb_hit r;
		const int2 h = hit[i];
		r.t = __int_as_float(h.x);
		r.nodes_visited = stats[i].x; r.prims_tested = stats[i].y; r.pad = 0;   // inner nodes visited, primitives tested
		if (h.y < 0) {
			r.t = FLT_MAX; r.prim = -1; r.object = -1; r.material = -1; r.front_face = 0; r.u = r.v = 0.0f;
			r.p[0] = r.p[1] = r.p[2] = 0.0f; r.n[0] = r.n[1] = r.n[2] = 0.0f;
		} else {
			const float4 fo = ro[i], fd = rd[i];
			const int2 info = __ldg(sv.prim_info + (h.y >> RTB_LEAF_TYPE_BITS));
			Surface s;
			reconstruct(sv, h.y, xyz(fo), xyz(fd), fo.w, r.t, true, s);
			r.prim = h.y >> RTB_LEAF_TYPE_BITS; r.object = info.y; r.material = info.x;
			r.p[0] = s.p.x; r.p[1] = s.p.y; r.p[2] = s.p.z;
			r.n[0] = s.n_shade.x; r.n[1] = s.n_shade.y; r.n[2] = s.n_shade.z;
			r.front_face = rt::dot(xyz(fd), s.n_geom) > 0.0f ? 0 : 1;
			r.u = s.u; r.v = s.v;
		}
		out[i] = r;
	}
}

// ------------------------------------------------------------------------------------------------
// host launch wrappers

int debug_bounds_report(unsigned long long* violations, unsigned long long* checks) {
#if RTB_DEBUG_BOUNDS
	if (cudaMemcpyFromSymbol(violations, g_bounds_violations, sizeof(unsigned long long) * RTB_BOUNDS_CLASSES) != cudaSuccess) return -1;
	if (cudaMemcpyFromSymbol(checks, g_bounds_checks, sizeof(unsigned long long)) != cudaSuccess) return -1;
	return 1;
#else
	(void)violations; (void)checks;
	return 0;
#endif
}

void query_occupancy(int device, LaunchCfg& lc) {
	cudaDeviceProp prop{};
	cudaGetDeviceProperties(&prop, device);
	int sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
	int occ_t = 0, occ_tm = 0, occ_s = 0, occ_g = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, traverse_kernel<0, STACK_SIZE>, TRAVERSE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tm, traverse_kernel<2, STACK_SIZE>, TRAVERSE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, shade_kernel, SHADE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, generate_kernel, STREAM_THREADS, 0);
	int occ_trav = occ_t < occ_tm ? occ_t : occ_tm;
	int occ_l = 0, occ_lm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, tail_kernel<false, STACK_SIZE>, TRAVERSE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lm, tail_kernel<true, STACK_SIZE>, TRAVERSE_THREADS, 0);
	int occ_tail = occ_l < occ_lm ? occ_l : occ_lm;
	lc.blocks_tail = sms * (occ_tail > 0 ? occ_tail : 1);
	lc.sms = sms;
	lc.blocks_traverse = sms * (occ_trav > 0 ? occ_trav : 1);
	lc.blocks_shade = sms * (occ_s > 0 ? occ_s : 1);
	lc.blocks_stream = sms * (occ_g > 0 ? occ_g : 1);
	int occ_b = 0;
	cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIN_COUNT_SMEM);
	cudaFuncSetAttribute(bin_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIN_PERMUTE_SMEM);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, bin_permute_kernel, BIN_THREADS, BIN_PERMUTE_SMEM);
	lc.blocks_bin = sms * (occ_b > 0 ? occ_b : 1);
}

void launch_generate(const BatchParams& bp, const rtb_camera& cam, const WaveView& wv, const LaunchCfg& lc, cudaStream_t st) {
	generate_kernel<<<lc.blocks_stream, STREAM_THREADS, 0, st>>>(bp, cam, wv);
}
void launch_traverse(const SceneView& sv, const BatchParams& bp, const WaveView& wv, uint32_t bounce, int q, const LaunchCfg& lc, cudaStream_t st) {
	// media need the per-path RNG inside traversal; scenes without media skip that code entirely
	// a walk keeps at most depth - 1 entries on its stack: 16 entries for shallow trees, 32 normally, 64 for the deep
	// trees a linear BVH over a large mesh can be (the flattener never hands over more than RTB_TREE_DEPTH_MAX levels)
	const bool small = sv.tree_depth <= 17, deep = sv.tree_depth > STACK_SIZE - 2;
#define RTB_LAUNCH_TRAVERSE(M) \
	do { if (small) traverse_kernel<M, 16><<<lc.blocks_traverse, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q); \
	     else if (deep) traverse_kernel<M, DEEP_STACK_SIZE><<<lc.blocks_traverse, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q); \
	     else traverse_kernel<M, STACK_SIZE><<<lc.blocks_traverse, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q); } while (0)
	if (sv.has_media == 0) RTB_LAUNCH_TRAVERSE(0);
	else if (sv.has_media == 1) RTB_LAUNCH_TRAVERSE(1);
	else RTB_LAUNCH_TRAVERSE(2);
#undef RTB_LAUNCH_TRAVERSE
}
void launch_tail(const SceneView& sv, const BatchParams& bp, const WaveView& wv, uint32_t bounce, int q, uint32_t threshold, const LaunchCfg& lc, cudaStream_t st) {
	const bool deep = sv.tree_depth > STACK_SIZE - 2;
	if (sv.has_media) {
		if (deep) tail_kernel<true, DEEP_STACK_SIZE><<<lc.blocks_tail, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q, threshold);
		else tail_kernel<true, STACK_SIZE><<<lc.blocks_tail, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q, threshold);
	} else {
		if (deep) tail_kernel<false, DEEP_STACK_SIZE><<<lc.blocks_tail, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q, threshold);
		else tail_kernel<false, STACK_SIZE><<<lc.blocks_tail, TRAVERSE_THREADS, 0, st>>>(sv, bp, wv, bounce, q, threshold);
	}
}
void launch_shade(const SceneView& sv, const BatchParams& bp, const WaveView& wv, uint32_t bounce, int q, const LaunchCfg& lc, cudaStream_t st) {
	shade_kernel<<<lc.blocks_shade, SHADE_THREADS, 0, st>>>(sv, bp, wv, bounce, q);
}
void launch_texture(const SceneView& sv, const WaveView& wv, uint32_t bounce, int q_out, const LaunchCfg& lc, cudaStream_t st) {
	const int blocks = lc.sms * 2 < lc.blocks_stream ? lc.sms * 2 : lc.blocks_stream;   // short work lists: a small grid keeps the launch cheap
	texture_kernel<<<blocks, STREAM_THREADS, 0, st>>>(sv, wv, bounce, q_out);
}
void launch_bin_rays(const SceneView& sv, const WaveView& wv, uint32_t bounce, int q_from, const LaunchCfg& lc, cudaStream_t st) {
	const uint32_t nbins = 1u << (3 * sv.bin_org_bits + 2 * sv.bin_dir_bits);
	bin_count_kernel<<<lc.blocks_bin, BIN_THREADS, BIN_COUNT_SMEM, st>>>(sv, wv, bounce, q_from);
	bin_scan_kernel<<<(nbins + BIN_SCAN_PER_BLOCK - 1) / BIN_SCAN_PER_BLOCK, BIN_SCAN_THREADS, 0, st>>>(wv, bounce);
	bin_permute_kernel<<<lc.blocks_bin, BIN_THREADS, BIN_PERMUTE_SMEM, st>>>(sv, wv, bounce, q_from);
}
void launch_accumulate(const BatchParams& bp, const WaveView& wv, float4* accum, float4* accum2, const LaunchCfg& lc, cudaStream_t st) {
	accumulate_kernel<<<lc.blocks_stream, STREAM_THREADS, 0, st>>>(bp, wv, accum, accum2);
	end_batch_kernel<<<1, 32, 0, st>>>(bp, wv);
}
void launch_resolve(const float4* accum, float4* out, uint32_t n, cudaStream_t st) {
	int blocks = (int)((n + STREAM_THREADS - 1) / STREAM_THREADS); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
	resolve_kernel<<<blocks, STREAM_THREADS, 0, st>>>(accum, out, n);
}
void launch_quantize(const float4* accum, uint8_t* rgb, uint32_t width, uint32_t height, int flip_rows, cudaStream_t st) {
	const uint32_t n = width * height;
	int blocks = (int)((n + STREAM_THREADS - 1) / STREAM_THREADS); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
	quantize_kernel<<<blocks, STREAM_THREADS, 0, st>>>(accum, rgb, width, height, flip_rows);
}
void launch_trace_rays(const SceneView& sv, const float4* ray_o, const float4* ray_d, uint32_t n, int2* hit_tmp, int2* stats_tmp,
                       rtb_hit* hits_out, uint32_t* work_counter, const LaunchCfg& lc, cudaStream_t st) {
	cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), st);
	if (sv.tree_depth > STACK_SIZE - 2) trace_rays_kernel<DEEP_STACK_SIZE><<<lc.blocks_traverse, TRAVERSE_THREADS, 0, st>>>(sv, ray_o, ray_d, n, hit_tmp, stats_tmp, work_counter);
	else trace_rays_kernel<STACK_SIZE><<<lc.blocks_traverse, TRAVERSE_THREADS, 0, st>>>(sv, ray_o, ray_d, n, hit_tmp, stats_tmp, work_counter);
	hit_record_kernel<<<lc.blocks_stream, STREAM_THREADS, 0, st>>>(sv, ray_o, ray_d, hit_tmp, stats_tmp, n, hits_out);
}

}  // namespace rtb

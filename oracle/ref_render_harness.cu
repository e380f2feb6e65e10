// oracle/ref_render_harness.cu — TEST INFRASTRUCTURE.  Drives the REFERENCE's own renderer — Renderer.cu,
// BVH.cu, SphereHittable.cu, Scenes.cu, cuHostRND.cpp compiled unmodified for sm_100a from where they lie
// under /root/reference — the way FirstApp::MakeApp / Run do (main/src/FirstApp.cpp:20-56,94-101), and dumps
// what the parity tests need.  Needs a GPU.
//
//   ref_render render <w> <h> <spp> <depth> <out.bin>   raw w*h float4 framebuffer (post clamp+sqrt) + JSON timing
//   ref_render scene <out.bin>                            the 488 spheres + raw material bytes SceneBook2BVH built
//   ref_render rng <out.bin>                              cuHostRND(512,1984) stream and device XORWOW streams
//   ref_render trace <rays.bin> <hits.bin>                the reference's own world->ClosestIntersection (BVH.cu:54-106 ->
//                                                         SphereHittable.cu:56-66,91-102) + getNormal (:43-50,75-83) on a ray
//                                                         file: n x rtb_ray (o.xyz, time, d.xyz, pad) in, n x 32-byte records
//                                                         (t, normal.xyz, point.xyz, sphere index or -1) out
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <glm/glm.hpp>

#include "Renderer.h"
#include "rt_engine/geometry/BVH.cuh"
#include "rt_engine/geometry/Scenes.h"
#include "rt_engine/geometry/SphereHittable.cuh"
#include "rt_engine/shaders/cu_Cameras.cuh"
#include "utilities/cuda_utilities/cuHostRND.h"

// Layout mirrors of classes whose members are private without accessors (Scenes.h:55-72,
// SphereHittable.cuh:104-158); used by the harness only.
struct SphereHandleView { aabb bounds; Material* material_ptr; Sphere* sphere_ptr; MovingSphere* moving_sphere_ptr; Hittable* hittable_ptr; };
struct SceneView { BVH_Handle* bvh; aabb* world_bounds; std::vector<SphereHandle> sphere_handles; };
static_assert(sizeof(SphereHandleView) == sizeof(SphereHandle), "SphereHandle layout");

__global__ void xorwow_kernel(unsigned long long seed, int n, float* out) {
	curandStateXORWOW_t st;
	curand_init(seed, 0, 0, &st);
	for (int i = 0; i < n; ++i) out[i] = curand_uniform(&st);
}

static int cmd_render(int argc, char** argv) {
	if (argc != 7) return 2;
	uint32_t w = atoi(argv[2]), h = atoi(argv[3]), spp = atoi(argv[4]), depth = atoi(argv[5]);
	auto cam = new MotionBlurCamera(glm::vec3(13, 2, 3), glm::vec3(0, 0, 0), glm::vec3(0, 1, 0), 30.0f, w / (float)h, 0.1f, 1.0f);
	SceneBook2BVH::Factory scene_factory{};
	SceneBook2BVH* scene = scene_factory.MakeScene();
	Renderer renderer = Renderer::MakeRenderer(w, h, spp, depth, cam, scene->getWorldPtr());
	std::vector<glm::vec4> fb((size_t)w * h);
	renderer.Render();   // synchronous (Renderer.cu:132-133); the first call also pays module load + clock ramp
	renderer.DownloadRenderbuffer(fb.data());   // the dumped image is the first render (fresh per-pixel RNG states)
	auto t0 = std::chrono::steady_clock::now();
	renderer.Render();   // timed: same work, RNG states carried on (Renderer.cu:191)
	auto t1 = std::chrono::steady_clock::now();
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
	FILE* o = fopen(argv[6], "wb"); if (!o) return 1;
	fwrite(fb.data(), sizeof(glm::vec4), fb.size(), o); fclose(o);
	double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	printf("REF_JSON {\"width\": %u, \"height\": %u, \"spp\": %u, \"depth\": %u, \"render_ms\": %.3f, \"mpaths_per_s\": %.3f}\n", w, h, spp, depth, ms,
	       (double)w * h * spp / ms * 1e-3);
	return 0;
}

static int cmd_scene(int argc, char** argv) {
	if (argc != 3) return 2;
	SceneBook2BVH::Factory scene_factory{};
	SceneBook2BVH* scene = scene_factory.MakeScene();
	const SceneView* view = reinterpret_cast<const SceneView*>(scene);
	FILE* o = fopen(argv[2], "wb"); if (!o) return 1;
	int32_t n = (int32_t)view->sphere_handles.size();
	fwrite(&n, 4, 1, o);
	for (int i = 0; i < n; ++i) {
		const SphereHandleView* hv = reinterpret_cast<const SphereHandleView*>(&view->sphere_handles[i]);
		float rec[16]; memset(rec, 0, sizeof rec);   // [0] moving flag, [1..7] geometry, [8..13] raw material bytes 8..32
		unsigned char mat[24];
		cudaMemcpy(mat, hv->material_ptr, 24, cudaMemcpyDeviceToHost);
		if (hv->moving_sphere_ptr) { rec[0] = 1.0f; cudaMemcpy(rec + 1, hv->moving_sphere_ptr, sizeof(MovingSphere), cudaMemcpyDeviceToHost); }
		else cudaMemcpy(rec + 1, hv->sphere_ptr, sizeof(Sphere), cudaMemcpyDeviceToHost);
		memcpy(rec + 8, mat + 8, 16);
		fwrite(rec, 4, 16, o);
	}
	fclose(o);
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// One thread per ray through the reference's device object graph, exactly as sample_world does for one segment
// (Renderer.cu:147-148): a fresh RayPayload, the world's virtual ClosestIntersection, then what a material would read
// back: the distance, the geometry's getNormal, and in_ray.at(rec.distance) (cu_materials.cuh:58,83).
struct TraceOut { float t, n[3], p[3]; int32_t index; const void* geom; };
__global__ void trace_kernel(const Hittable* world, const float* rays, int n, TraceOut* out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* r = rays + 8 * (size_t)i;
	Ray ray(glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]);
	RayPayload rec{};
	bool hit = world->ClosestIntersection(ray, rec);
	TraceOut o{}; o.t = rec.distance; o.index = -1; o.geom = nullptr;
	if (hit) {
		glm::vec3 nn = Sphere::getNormal(ray, rec);   // Sphere:: and MovingSphere::TraceRecord share one layout {ptr, normal}
		glm::vec3 pp = ray.at(rec.distance);
		o.n[0] = nn.x; o.n[1] = nn.y; o.n[2] = nn.z; o.p[0] = pp.x; o.p[1] = pp.y; o.p[2] = pp.z;
		o.geom = reinterpret_cast<const Sphere::TraceRecord*>(&rec.payload)->sphere;
	}
	out[i] = o;
}

static int cmd_trace(int argc, char** argv) {
	if (argc != 4) return 2;
	FILE* f = fopen(argv[2], "rb"); if (!f) { perror("rays"); return 1; }
	fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
	int n = (int)(bytes / 32);
	std::vector<float> rays(8 * (size_t)n);
	if (fread(rays.data(), 32, n, f) != (size_t)n) return 1;
	fclose(f);
	SceneBook2BVH::Factory scene_factory{};
	SceneBook2BVH* scene = scene_factory.MakeScene();
	const SceneView* view = reinterpret_cast<const SceneView*>(scene);
	cudaDeviceSetLimit(cudaLimitStackSize, 8192);   // as Renderer::Render does for the virtual calls (Renderer.cu:124)
	float* d_rays; TraceOut* d_out;
	cudaMalloc(&d_rays, rays.size() * 4); cudaMalloc(&d_out, sizeof(TraceOut) * (size_t)n);
	cudaMemcpy(d_rays, rays.data(), rays.size() * 4, cudaMemcpyHostToDevice);
	trace_kernel<<<(n + 63) / 64, 64>>>(scene->getWorldPtr(), d_rays, n, d_out);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
	std::vector<TraceOut> out(n);
	cudaMemcpy(out.data(), d_out, sizeof(TraceOut) * (size_t)n, cudaMemcpyDeviceToHost);
	// geometry pointer -> index of the sphere in construction order (SceneBook2BVH::sphere_handles)
	std::vector<std::pair<const void*, int>> ptrs;
	for (int i = 0; i < (int)view->sphere_handles.size(); ++i) {
		const SphereHandleView* hv = reinterpret_cast<const SphereHandleView*>(&view->sphere_handles[i]);
		ptrs.push_back({hv->moving_sphere_ptr ? (const void*)hv->moving_sphere_ptr : (const void*)hv->sphere_ptr, i});
	}
	std::sort(ptrs.begin(), ptrs.end());
	FILE* o = fopen(argv[3], "wb"); if (!o) return 1;
	int hits = 0;
	for (int i = 0; i < n; ++i) {
		if (out[i].geom) {
			auto it = std::lower_bound(ptrs.begin(), ptrs.end(), std::make_pair(out[i].geom, -1));
			if (it == ptrs.end() || it->first != out[i].geom) { fprintf(stderr, "unknown geometry pointer\n"); return 1; }
			out[i].index = it->second; ++hits;
		}
		fwrite(&out[i], 32, 1, o);   // (t, n, p, index): the first 32 bytes of the record
	}
	fclose(o);
	printf("REF_TRACE {\"rays\": %d, \"hits\": %d}\n", n, hits);
	return 0;
}

static int cmd_rng(int argc, char** argv) {
	if (argc != 3) return 2;
	FILE* o = fopen(argv[2], "wb"); if (!o) return 1;
	{   // the reference's scene-construction stream (device generator behind the cuRAND host API)
		cuHostRND rnd(512, 1984);
		std::vector<float> u(4608);
		for (auto& x : u) x = rnd.next();
		fwrite(u.data(), 4, u.size(), o);
	}
	const unsigned long long seeds[4] = {1984ull, 1985ull, 1984ull + 45000ull, 1984ull + 89999ull};
	float* d; cudaMalloc(&d, 64 * 4);
	for (int s = 0; s < 4; ++s) {
		xorwow_kernel<<<1, 1>>>(seeds[s], 64, d);
		float hbuf[64]; cudaMemcpy(hbuf, d, sizeof hbuf, cudaMemcpyDeviceToHost);
		fwrite(hbuf, 4, 64, o);
	}
	fclose(o);
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: ref_render render|scene|rng|trace ...\n"); return 2; }
	if (!strcmp(argv[1], "render")) return cmd_render(argc, argv);
	if (!strcmp(argv[1], "scene")) return cmd_scene(argc, argv);
	if (!strcmp(argv[1], "rng")) return cmd_rng(argc, argv);
	if (!strcmp(argv[1], "trace")) return cmd_trace(argc, argv);
	return 2;
}

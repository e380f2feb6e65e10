// oracle/ref_sphere_index_harness.cu — TEST INFRASTRUCTURE.  The ground truth of the reference's only unit test
// (google_testing/test.cpp, SphereTest.DeviceSphereIndexTest) computed by the REFERENCE's own functions, compiled
// from the reference's headers where they lie under /root/reference:
//   _sphere_closest_intersection   main/src/rt_engine/geometry/SphereHittable.cuh:15-33  (__host__ __device__ inline)
//   PinholeCamera ctor, sample_ray  main/src/rt_engine/shaders/cu_Cameras.cuh:12-31
// in the recipe of test.cpp:87-106 (host) and, with a GPU, test.cpp:112-135 (device).  gtest itself is a NuGet
// package that is not available here, so the recipe is driven from this harness and asserted from pytest.
//
//   ref_sphere_index host   <spheres.bin> <w> <h> <out.bin>     runs on the CPU (no GPU needed)
//   ref_sphere_index device <spheres.bin> <w> <h> <out.bin>     the same through a kernel (needs a GPU)
//   spheres.bin: n x (center.xyz, radius) float32
//   out.bin    : 12 floats camera (o, u, v, w), then w*h x int32 closest-sphere index (-1 = none), then
//                w*h x 32-byte rtb_ray records (o.xyz, time 0, d.xyz, 0): the rays the reference built, so that
//                the implementation under test traces the same rays.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <glm/glm.hpp>

#include "rt_engine/ray_data.cuh"
#include "rt_engine/shaders/cu_Cameras.cuh"
#include "rt_engine/geometry/SphereHittable.cuh"

// Pixel -> NDC of the test (u = x/(W-1)*2-1, not the renderer's pixel centres), then the reference's camera.
__host__ __device__ inline Ray pixel_ray(const PinholeCamera& cam, int x, int y, int w, int h) {
	float u = x / (w - 1.0f) * 2 - 1;
	float v = y / (h - 1.0f) * 2 - 1;
	return cam.sample_ray(u, v);
}

// First sphere, in list order, with the smallest distance below _MISS_DIST.
__host__ __device__ inline int closest_sphere(const Ray& ray, const Sphere* sp, int n) {
	RayPayload rec{};
	int best = -1;
	for (int i = 0; i < n; i++) {
		float dist = _sphere_closest_intersection(ray, sp[i].center, sp[i].radius);
		if (dist < rec.distance) { best = i; rec.distance = dist; }
	}
	return best;
}

__global__ void index_kernel(const Sphere* sp, int n, PinholeCamera cam, int w, int h, int* out, float* rays) {
	int x = blockDim.x * blockIdx.x + threadIdx.x, y = blockDim.y * blockIdx.y + threadIdx.y;
	if (x >= w || y >= h) return;
	Ray ray = pixel_ray(cam, x, y, w, h);
	size_t gid = (size_t)y * w + x;
	out[gid] = closest_sphere(ray, sp, n);
	float* r = rays + 8 * gid;
	r[0] = ray.o.x; r[1] = ray.o.y; r[2] = ray.o.z; r[3] = 0.0f; r[4] = ray.d.x; r[5] = ray.d.y; r[6] = ray.d.z; r[7] = 0.0f;
}

int main(int argc, char** argv) {
	if (argc != 6) { fprintf(stderr, "usage: ref_sphere_index host|device spheres.bin w h out.bin\n"); return 2; }
	const bool device = !strcmp(argv[1], "device");
	FILE* f = fopen(argv[2], "rb"); if (!f) { perror("spheres"); return 1; }
	fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
	int n = (int)(bytes / 16);
	std::vector<float> raw(4 * (size_t)n);
	if (fread(raw.data(), 16, n, f) != (size_t)n) return 1;
	fclose(f);
	std::vector<Sphere> spheres;
	for (int i = 0; i < n; ++i) spheres.push_back(Sphere(glm::vec3(raw[4 * i], raw[4 * i + 1], raw[4 * i + 2]), raw[4 * i + 3]));
	int w = atoi(argv[3]), h = atoi(argv[4]);
	// SphereTest::SetUp  test.cpp:21-26
	PinholeCamera cam(glm::vec3(0, 1, -4), glm::vec3(0, 1, 0), glm::vec3(0, 1, 0), 90.0f, w / (float)h);
	std::vector<int32_t> idx((size_t)w * h);
	std::vector<float> rays(8 * (size_t)w * h, 0.0f);
	if (!device) {
		for (int y = 0; y < h; y++)
			for (int x = 0; x < w; x++) {
				Ray ray = pixel_ray(cam, x, y, w, h);
				size_t gid = (size_t)y * w + x;
				idx[gid] = closest_sphere(ray, spheres.data(), n);
				float* r = rays.data() + 8 * gid;
				r[0] = ray.o.x; r[1] = ray.o.y; r[2] = ray.o.z; r[4] = ray.d.x; r[5] = ray.d.y; r[6] = ray.d.z;
			}
	} else {
		Sphere* d_sp; int* d_idx; float* d_rays;
		cudaMalloc(&d_sp, sizeof(Sphere) * n); cudaMalloc(&d_idx, 4 * idx.size()); cudaMalloc(&d_rays, 4 * rays.size());
		cudaMemcpy(d_sp, spheres.data(), sizeof(Sphere) * n, cudaMemcpyHostToDevice);
		dim3 threads(8, 8, 1), blocks((w + 7) / 8, (h + 7) / 8, 1);
		index_kernel<<<blocks, threads>>>(d_sp, n, cam, w, h, d_idx, d_rays);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
		cudaMemcpy(idx.data(), d_idx, 4 * idx.size(), cudaMemcpyDeviceToHost);
		cudaMemcpy(rays.data(), d_rays, 4 * rays.size(), cudaMemcpyDeviceToHost);
	}
	FILE* o = fopen(argv[5], "wb"); if (!o) { perror("out"); return 1; }
	float camf[12] = {cam.o.x, cam.o.y, cam.o.z, cam.u.x, cam.u.y, cam.u.z, cam.v.x, cam.v.y, cam.v.z, cam.w.x, cam.w.y, cam.w.z};
	fwrite(camf, 4, 12, o);
	fwrite(idx.data(), 4, idx.size(), o);
	fwrite(rays.data(), 4, rays.size(), o);
	fclose(o);
	return 0;
}

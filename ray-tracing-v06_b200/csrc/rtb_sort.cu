// rtb_sort.cu — EXPERIMENT (RTB_SORT_EXPERIMENT=<mode>, profiling renders only): globally reorders the live ray queue
// between shade(b) and traverse(b+1) by a coherence key, to measure what lane coherence is worth to the traverse kernel
// before investing in an in-graph binning pass.  The sort itself is cub::DeviceRadixSort (library code) and is NOT part of
// the timed classes; only the traverse time on the reordered queue is read.  See profiles/r2_experiments.md.
#include <cub/device/device_radix_sort.cuh>

#include "rtb_renderer.h"

namespace rtb {

struct SortBounds { float mn[3], inv[3]; };

__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // 10 bits -> every third bit
	v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
	return v;
}

// mode bits: 0-3 direction bits per octahedral axis (0 = none, else 2^k x 2^k cells after the octant), 4-7 origin bits per axis
__global__ void sort_keys_kernel(const float4* __restrict__ ro, const float4* __restrict__ rd, const uint32_t* n_ptr, uint32_t* keys, uint32_t* idx,
                                 int dir_bits, int org_bits, int dir_major, SortBounds sb) {
	const uint32_t n = *n_ptr;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float4 o = ro[i], d = rd[i];
		uint32_t kd = 0, ko = 0;
		if (dir_bits > 0) {
			// octahedral map of the direction to [0,1]^2, quantised to 2^dir_bits cells per axis
			const float s = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
			float u = d.x * s, v = d.z * s;
			if (d.y < 0.0f) { const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f); u = uu; v = vv; }
			const int q = 1 << dir_bits;
			int iu = (int)((u * 0.5f + 0.5f) * q), iv = (int)((v * 0.5f + 0.5f) * q);
			iu = min(max(iu, 0), q - 1); iv = min(max(iv, 0), q - 1);
			kd = (uint32_t)(iv * q + iu);
		}
		if (org_bits > 0) {
			const int q = 1 << org_bits;
			int ix = (int)((o.x - sb.mn[0]) * sb.inv[0] * q), iy = (int)((o.y - sb.mn[1]) * sb.inv[1] * q), iz = (int)((o.z - sb.mn[2]) * sb.inv[2] * q);
			ix = min(max(ix, 0), q - 1); iy = min(max(iy, 0), q - 1); iz = min(max(iz, 0), q - 1);
			ko = spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2);
		}
		keys[i] = dir_major ? ((kd << (3 * org_bits)) | ko) : ((ko << (2 * dir_bits)) | kd);
		idx[i] = i;
	}
}

__global__ void permute_kernel(const uint32_t* __restrict__ idx, const uint32_t* n_ptr, const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a2,
                               float4* __restrict__ b0, float4* __restrict__ b1, float4* __restrict__ b2) {
	const uint32_t n = *n_ptr;
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const uint32_t i = idx[k];
		b0[k] = a0[i]; b1[k] = a1[i]; b2[k] = a2[i];
	}
}

struct SortScratch { uint32_t *keys = nullptr, *keys2 = nullptr, *idx = nullptr, *idx2 = nullptr; float4 *b0 = nullptr, *b1 = nullptr, *b2 = nullptr; void* temp = nullptr; size_t temp_bytes = 0, cap = 0; };
static SortScratch g_scratch;

// Reorders queue `q` (0/1) of bounce `bounce`: synchronises, so only for profiling renders.
int experimental_sort_queue(rtb_renderer* r, uint32_t bounce, int q, int mode, const float* world_min, const float* world_max) {
	cudaStream_t st = r->stream;
	uint32_t n = 0;
	CUDA_TRY(cudaMemcpyAsync(&n, r->wv.n_live + bounce, 4, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (n < 2) return RTB_OK;
	SortScratch& s = g_scratch;
	if (s.cap < r->wave_paths) {
		cudaFree(s.keys); cudaFree(s.keys2); cudaFree(s.idx); cudaFree(s.idx2); cudaFree(s.b0); cudaFree(s.b1); cudaFree(s.b2); cudaFree(s.temp);
		const size_t P = r->wave_paths;
		CUDA_TRY(cudaMalloc(&s.keys, P * 4)); CUDA_TRY(cudaMalloc(&s.keys2, P * 4)); CUDA_TRY(cudaMalloc(&s.idx, P * 4)); CUDA_TRY(cudaMalloc(&s.idx2, P * 4));
		CUDA_TRY(cudaMalloc(&s.b0, P * 16)); CUDA_TRY(cudaMalloc(&s.b1, P * 16)); CUDA_TRY(cudaMalloc(&s.b2, P * 16));
		s.temp_bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, s.temp_bytes, s.keys, s.keys2, s.idx, s.idx2, (int)P, 0, 32, st);
		CUDA_TRY(cudaMalloc(&s.temp, s.temp_bytes));
		s.cap = P;
	}
	const int dir_bits = mode & 15, org_bits = (mode >> 4) & 15, dir_major = (mode >> 8) & 1;
	SortBounds sb;
	for (int k = 0; k < 3; ++k) { sb.mn[k] = world_min[k]; const float e = world_max[k] - world_min[k]; sb.inv[k] = e > 0.0f ? 1.0f / e : 0.0f; }
	sort_keys_kernel<<<148 * 8, 256, 0, st>>>(r->wv.ray_o[q], r->wv.ray_d[q], r->wv.n_live + bounce, s.keys, s.idx, dir_bits, org_bits, dir_major, sb);
	const int bits = 2 * dir_bits + 3 * org_bits;
	size_t tb = s.temp_bytes;
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(s.temp, tb, s.keys, s.keys2, s.idx, s.idx2, (int)n, 0, bits > 0 ? bits : 1, st));
	permute_kernel<<<148 * 8, 256, 0, st>>>(s.idx2, r->wv.n_live + bounce, r->wv.ray_o[q], r->wv.ray_d[q], r->wv.thr[q], s.b0, s.b1, s.b2);
	CUDA_TRY(cudaMemcpyAsync(r->wv.ray_o[q], s.b0, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(r->wv.ray_d[q], s.b1, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(r->wv.thr[q], s.b2, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaGetLastError());
	return RTB_OK;
}

}  // namespace rtb

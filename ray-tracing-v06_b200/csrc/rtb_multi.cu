// rtb_multi.cu — several GPUs of one box behind the C ABI (include/rtb.h, "several GPUs of one box").
//
// The reference renders on one device from one host thread (main/src/FirstApp.cpp:39-40,94-101).  Here one process
// drives one rtb_renderer per device: the scene is flattened once and its arena uploaded to every device; a render's
// sample range is split into contiguous per-device sub-ranges (counter-based random streams make those the very samples
// a single device would draw); every device renders on its own stream, concurrently; then the per-device radiance sums
// are added onto the first device - one ncclReduce(sum) over NVLink per render (NCCL is loaded with dlopen, so a
// single-GPU user never needs it), or one kernel on the first device that reads its peers' accumulators through peer
// memory and adds them in device order.  Host-side work per device (graph capture, uploads) runs on one thread per device.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rtb_renderer.h"

using namespace rtb;

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (the handful of entry points a single-process reduce needs; signatures of nccl.h 2.x)

namespace {

typedef struct ncclComm* ncclComm_t;
enum { NCCL_SUCCESS = 0, NCCL_FLOAT = 7, NCCL_SUM = 0 };

struct NcclApi {
	void* lib = nullptr;
	int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	bool ok() const { return CommInitAll && CommDestroy && Reduce && GroupStart && GroupEnd && GetErrorString; }
};

NcclApi load_nccl() {
	NcclApi a;
	const char* names[] = {getenv("RTB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char* n : names) {
		if (!n || !*n) continue;
		a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (a.lib) break;
	}
	if (!a.lib) return a;
	a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(a.lib, "ncclCommInitAll"));
	a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
	a.Reduce = reinterpret_cast<decltype(a.Reduce)>(dlsym(a.lib, "ncclReduce"));
	a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.lib, "ncclGroupStart"));
	a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.lib, "ncclGroupEnd"));
	a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
	return a;
}

// total[i] = a[0][i] + a[1][i] + ... in device order: the first device reads its peers' accumulators over NVLink.
struct PeerList { const float4* a[16]; int n; };
__global__ void __launch_bounds__(256)
reduce_peers_kernel(PeerList peers, float4* __restrict__ total, uint32_t n) {
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		float4 s = peers.a[0][i];
		for (int k = 1; k < peers.n; ++k) {
			const float4 v = __ldcs(peers.a[k] + i);
			s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
		}
		total[i] = s;
	}
}

}  // namespace

struct rtb_multi_renderer {
	std::vector<rtb_renderer*> r;          // one per device, r[0] is the root of the reduction
	std::vector<cudaStream_t> stream;      // the stream each device's render + reduce is ordered on
	std::vector<cudaEvent_t> done;         // "this device's accumulators are final" (peer-memory reduce)
	std::vector<cudaEvent_t> t0, t1;       // device time of the last render, per device
	int mode = RTB_REDUCE_P2P;
	NcclApi nccl;
	std::vector<ncclComm_t> comm;
	float4 *d_total = nullptr, *d_total2 = nullptr;   // cross-device sums on device r[0]
	uint32_t width = 0, height = 0;
	bool rendered = false, variance = false;
	uint64_t reduce_launches = 0;
};

namespace {

// Runs fn(i) for every device slot on its own host thread and returns the first failure (with its message).
template <class F>
int for_each_device(rtb_multi_renderer* m, F fn) {
	const int n = (int)m->r.size();
	std::vector<int> rc(n, RTB_OK);
	std::vector<std::string> msg(n);
	auto body = [&](int i) { rc[i] = fn(i); if (rc[i] != RTB_OK) msg[i] = rtb_last_error(); };
	if (n == 1) body(0);
	else {
		std::vector<std::thread> th;
		for (int i = 0; i < n; ++i) th.emplace_back(body, i);
		for (auto& t : th) t.join();
	}
	for (int i = 0; i < n; ++i) if (rc[i] != RTB_OK) return fail(rc[i], "device slot " + std::to_string(i) + ": " + msg[i]);
	return RTB_OK;
}

int ensure_totals(rtb_multi_renderer* m, uint32_t w, uint32_t h) {
	if (m->d_total && m->width == w && m->height == h) return RTB_OK;
	CUDA_TRY(cudaSetDevice(m->r[0]->device));
	CUDA_TRY(cudaStreamSynchronize(m->stream[0]));
	cudaFree(m->d_total); cudaFree(m->d_total2); m->d_total = m->d_total2 = nullptr;
	const size_t bytes = (size_t)w * h * sizeof(float4);
	CUDA_TRY(cudaMalloc(&m->d_total, bytes));
	CUDA_TRY(cudaMalloc(&m->d_total2, bytes));
	CUDA_TRY(cudaMemset(m->d_total, 0, bytes));
	CUDA_TRY(cudaMemset(m->d_total2, 0, bytes));
	m->width = w; m->height = h;
	return RTB_OK;
}

}  // namespace

extern "C" {

int rtb_multi_renderer_create(rtb_multi_renderer** out, const int* devices, int n, int reduce_mode) {
	if (!out || n < 1 || n > 16) return fail(RTB_ERR_INVALID, "rtb_multi_renderer_create: need 1..16 devices");
	if (reduce_mode < RTB_REDUCE_AUTO || reduce_mode > RTB_REDUCE_P2P) return fail(RTB_ERR_INVALID, "rtb_multi_renderer_create: bad reduce mode");
	if (const char* e = getenv("RTB_MULTI_REDUCE")) {   // diagnostics switch: nccl | p2p
		if (!strcmp(e, "nccl")) reduce_mode = RTB_REDUCE_NCCL; else if (!strcmp(e, "p2p")) reduce_mode = RTB_REDUCE_P2P;
	}
	rtb_multi_renderer* m = new rtb_multi_renderer();
	std::vector<int> dev(n);
	for (int i = 0; i < n; ++i) {
		dev[i] = devices ? devices[i] : i;
		for (int k = 0; k < i; ++k) if (dev[k] == dev[i]) { rtb_multi_renderer_destroy(m); return fail(RTB_ERR_INVALID, "rtb_multi_renderer_create: a device is listed twice"); }
	}
	auto bail = [&](int rc) { const std::string keep = rtb_last_error(); rtb_multi_renderer_destroy(m); set_error(keep); return rc; };
	for (int i = 0; i < n; ++i) {
		rtb_renderer* r = nullptr;
		int rc = rtb_renderer_create(&r, dev[i]);
		if (rc) return bail(rc);
		m->r.push_back(r);
		cudaStream_t s = nullptr; cudaEvent_t e = nullptr, a = nullptr, b = nullptr;
		if (cudaSetDevice(dev[i]) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
			fail(RTB_ERR_CUDA, std::string("rtb_multi_renderer_create: ") + cudaGetErrorString(cudaGetLastError()));
			return bail(RTB_ERR_CUDA);
		}
		m->stream.push_back(s); m->done.push_back(e); m->t0.push_back(a); m->t1.push_back(b);
	}
	// how the sums travel: NCCL when asked for or available, else peer memory
	if (n > 1 && reduce_mode != RTB_REDUCE_P2P) {
		m->nccl = load_nccl();
		if (m->nccl.ok()) {
			m->comm.assign(n, nullptr);
			const int rc = m->nccl.CommInitAll(m->comm.data(), n, dev.data());
			if (rc == NCCL_SUCCESS) m->mode = RTB_REDUCE_NCCL;
			else {
				m->comm.clear();
				if (reduce_mode == RTB_REDUCE_NCCL) { fail(RTB_ERR_CUDA, std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(rc)); return bail(RTB_ERR_CUDA); }
			}
		} else if (reduce_mode == RTB_REDUCE_NCCL) {
			fail(RTB_ERR_UNSUPPORTED, "rtb_multi_renderer_create: libnccl.so.2 could not be loaded (set RTB_NCCL_LIB, or use RTB_REDUCE_P2P)");
			return bail(RTB_ERR_UNSUPPORTED);
		}
	}
	if (n > 1 && m->mode == RTB_REDUCE_P2P) {
		cudaSetDevice(dev[0]);
		for (int i = 1; i < n; ++i) {
			int can = 0;
			cudaDeviceCanAccessPeer(&can, dev[0], dev[i]);
			if (!can) { fail(RTB_ERR_UNSUPPORTED, "rtb_multi_renderer_create: device " + std::to_string(dev[0]) + " cannot read device " + std::to_string(dev[i]) + " (no peer access)"); return bail(RTB_ERR_UNSUPPORTED); }
			cudaError_t e = cudaDeviceEnablePeerAccess(dev[i], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { fail(RTB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return bail(RTB_ERR_CUDA); }
			cudaGetLastError();
		}
	}
	*out = m;
	return RTB_OK;
}

void rtb_multi_renderer_destroy(rtb_multi_renderer* m) {
	if (!m) return;
	for (size_t i = 0; i < m->r.size(); ++i) {
		cudaSetDevice(m->r[i]->device);
		if (i < m->stream.size() && m->stream[i]) cudaStreamSynchronize(m->stream[i]);
	}
	for (ncclComm_t c : m->comm) if (c) m->nccl.CommDestroy(c);
	if (!m->r.empty()) { cudaSetDevice(m->r[0]->device); cudaFree(m->d_total); cudaFree(m->d_total2); }
	for (size_t i = 0; i < m->r.size(); ++i) {
		cudaSetDevice(m->r[i]->device);
		if (i < m->stream.size()) { cudaEventDestroy(m->done[i]); cudaEventDestroy(m->t0[i]); cudaEventDestroy(m->t1[i]); cudaStreamDestroy(m->stream[i]); }
		rtb_renderer_destroy(m->r[i]);
	}
	cudaGetLastError();
	delete m;
}

int rtb_multi_device_count(const rtb_multi_renderer* m) { return m ? (int)m->r.size() : 0; }
rtb_renderer* rtb_multi_renderer_get(rtb_multi_renderer* m, int i) { return (m && i >= 0 && i < (int)m->r.size()) ? m->r[i] : nullptr; }
int rtb_multi_reduce_mode(const rtb_multi_renderer* m) { return m ? m->mode : RTB_ERR_INVALID; }

int rtb_multi_set_scene(rtb_multi_renderer* m, rtb_scene* s) {
	if (!m || !s) return fail(RTB_ERR_INVALID, "rtb_multi_set_scene: null argument");
	int rc = rtb_renderer_set_scene(m->r[0], s);          // ONE flatten (and BVH build) ...
	if (rc) return rc;
	return for_each_device(m, [&](int i) { return i == 0 ? RTB_OK : rtb_renderer_share_scene(m->r[i], m->r[0]); });   // ... N uploads
}

int rtb_multi_set_camera(rtb_multi_renderer* m, const rtb_camera* cam) {
	if (!m || !cam) return fail(RTB_ERR_INVALID, "rtb_multi_set_camera: null argument");
	for (rtb_renderer* r : m->r) { int rc = rtb_renderer_set_camera(r, cam); if (rc) return rc; }
	return RTB_OK;
}

int rtb_multi_render(rtb_multi_renderer* m, const rtb_render_params* p) {
	if (!m || !p) return fail(RTB_ERR_INVALID, "rtb_multi_render: null argument");
	if (p->sample_end < p->sample_begin) return fail(RTB_ERR_INVALID, "rtb_multi_render: sample_end < sample_begin");
	if (p->width == 0 || p->height == 0) return fail(RTB_ERR_INVALID, "rtb_multi_render: zero width/height");
	const int n = (int)m->r.size();
	int rc = ensure_totals(m, p->width, p->height);
	if (rc) return rc;
	const uint64_t spp = p->sample_end - p->sample_begin;
	// every device renders its sample sub-range on its own stream
	rc = for_each_device(m, [&](int i) {
		rtb_render_params q = *p;
		q.sample_begin = p->sample_begin + (uint32_t)(spp * i / n);
		q.sample_end = p->sample_begin + (uint32_t)(spp * (i + 1) / n);
		CUDA_TRY(cudaSetDevice(m->r[i]->device));
		CUDA_TRY(cudaEventRecord(m->t0[i], m->stream[i]));
		return rtb_render(m->r[i], &q, m->stream[i]);      // (an empty sub-range still sizes / clears this device's accumulators)
	});
	if (rc) return rc;
	m->variance = (p->flags & RTB_RENDER_VARIANCE) != 0;
	const size_t count = (size_t)p->width * p->height * 4;
	if (n == 1) {
		CUDA_TRY(cudaSetDevice(m->r[0]->device));
		CUDA_TRY(cudaMemcpyAsync(m->d_total, m->r[0]->d_accum, count * 4, cudaMemcpyDeviceToDevice, m->stream[0]));
		if (m->variance) CUDA_TRY(cudaMemcpyAsync(m->d_total2, m->r[0]->d_accum2, count * 4, cudaMemcpyDeviceToDevice, m->stream[0]));
	} else if (m->mode == RTB_REDUCE_NCCL) {
		for (int pass = 0; pass < (m->variance ? 2 : 1); ++pass) {
			int e = m->nccl.GroupStart();
			for (int i = 0; i < n && e == NCCL_SUCCESS; ++i)
				e = m->nccl.Reduce(pass ? m->r[i]->d_accum2 : m->r[i]->d_accum, pass ? m->d_total2 : m->d_total, count, NCCL_FLOAT, NCCL_SUM, 0, m->comm[i], m->stream[i]);
			const int e2 = m->nccl.GroupEnd();
			if (e == NCCL_SUCCESS) e = e2;
			if (e != NCCL_SUCCESS) return fail(RTB_ERR_CUDA, std::string("ncclReduce: ") + m->nccl.GetErrorString(e));
			m->reduce_launches += n;
		}
	} else {
		for (int i = 1; i < n; ++i) {
			CUDA_TRY(cudaSetDevice(m->r[i]->device));
			CUDA_TRY(cudaEventRecord(m->done[i], m->stream[i]));
		}
		CUDA_TRY(cudaSetDevice(m->r[0]->device));
		for (int i = 1; i < n; ++i) CUDA_TRY(cudaStreamWaitEvent(m->stream[0], m->done[i], 0));
		const uint32_t npx = p->width * p->height;
		int blocks = (int)((npx + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
		for (int pass = 0; pass < (m->variance ? 2 : 1); ++pass) {
			PeerList pl{}; pl.n = n;
			for (int i = 0; i < n; ++i) pl.a[i] = pass ? m->r[i]->d_accum2 : m->r[i]->d_accum;
			reduce_peers_kernel<<<blocks, 256, 0, m->stream[0]>>>(pl, pass ? m->d_total2 : m->d_total, npx);
			m->reduce_launches += 1;
		}
		CUDA_TRY(cudaGetLastError());
		// the peers' accumulators must not be rendered into again before the root has read them
		CUDA_TRY(cudaEventRecord(m->done[0], m->stream[0]));
		for (int i = 1; i < n; ++i) { CUDA_TRY(cudaSetDevice(m->r[i]->device)); CUDA_TRY(cudaStreamWaitEvent(m->stream[i], m->done[0], 0)); }
	}
	for (int i = 0; i < n; ++i) { CUDA_TRY(cudaSetDevice(m->r[i]->device)); CUDA_TRY(cudaEventRecord(m->t1[i], m->stream[i])); }
	m->rendered = true;
	return RTB_OK;
}

int rtb_multi_synchronize(rtb_multi_renderer* m) {
	if (!m) return fail(RTB_ERR_INVALID, "rtb_multi_synchronize: null renderer");
	for (size_t i = 0; i < m->r.size(); ++i) { CUDA_TRY(cudaSetDevice(m->r[i]->device)); CUDA_TRY(cudaStreamSynchronize(m->stream[i])); }
	return RTB_OK;
}

int rtb_multi_download(rtb_multi_renderer* m, float* host_rgba) {
	if (!m || !host_rgba) return fail(RTB_ERR_INVALID, "rtb_multi_download: null argument");
	if (!m->rendered) return fail(RTB_ERR_STATE, "rtb_multi_download: nothing rendered");
	rtb_renderer* r = m->r[0];
	int rc = rtb_resolve_from(r, m->d_total, nullptr, m->stream[0]);
	if (rc) return rc;
	CUDA_TRY(cudaMemcpyAsync(host_rgba, r->d_out, (size_t)m->width * m->height * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
	CUDA_TRY(cudaStreamSynchronize(r->stream));
	return RTB_OK;
}

int rtb_multi_download_accum(rtb_multi_renderer* m, float* host_sum, float* host_sum2) {
	if (!m) return fail(RTB_ERR_INVALID, "rtb_multi_download_accum: null renderer");
	if (!m->rendered) return fail(RTB_ERR_STATE, "rtb_multi_download_accum: nothing rendered");
	if (host_sum2 && !m->variance) return fail(RTB_ERR_STATE, "rtb_multi_download_accum: the last render did not ask for RTB_RENDER_VARIANCE");
	CUDA_TRY(cudaSetDevice(m->r[0]->device));
	const size_t bytes = (size_t)m->width * m->height * sizeof(float4);
	if (host_sum) CUDA_TRY(cudaMemcpyAsync(host_sum, m->d_total, bytes, cudaMemcpyDeviceToHost, m->stream[0]));
	if (host_sum2) CUDA_TRY(cudaMemcpyAsync(host_sum2, m->d_total2, bytes, cudaMemcpyDeviceToHost, m->stream[0]));
	CUDA_TRY(cudaStreamSynchronize(m->stream[0]));
	return RTB_OK;
}

int rtb_multi_download_rgb8(rtb_multi_renderer* m, uint8_t* host_rgb, int flip_rows) {
	if (!m || !host_rgb) return fail(RTB_ERR_INVALID, "rtb_multi_download_rgb8: null argument");
	if (!m->rendered) return fail(RTB_ERR_STATE, "rtb_multi_download_rgb8: nothing rendered");
	CUDA_TRY(cudaSetDevice(m->r[0]->device));
	CUDA_TRY(cudaStreamSynchronize(m->stream[0]));
	return rtb_quantize_from(m->r[0], m->d_total, host_rgb, flip_rows);
}

int rtb_multi_get_counters(rtb_multi_renderer* m, rtb_counters* out) {
	if (!m || !out) return fail(RTB_ERR_INVALID, "rtb_multi_get_counters: null argument");
	memset(out, 0, sizeof *out);
	for (size_t i = 0; i < m->r.size(); ++i) {
		rtb_counters c{};
		int rc = rtb_get_counters(m->r[i], &c);
		if (rc) return rc;
		out->paths += c.paths; out->rays += c.rays; out->launches += c.launches; out->batches += c.batches;
		if (m->rendered) {
			CUDA_TRY(cudaSetDevice(m->r[i]->device));
			CUDA_TRY(cudaStreamSynchronize(m->stream[i]));
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, m->t0[i], m->t1[i]) == cudaSuccess) { if (ms > out->render_ms) out->render_ms = ms; } else cudaGetLastError();
		}
	}
	out->launches += m->reduce_launches;
	return RTB_OK;
}

int rtb_multi_reset_counters(rtb_multi_renderer* m) {
	if (!m) return fail(RTB_ERR_INVALID, "rtb_multi_reset_counters: null renderer");
	for (rtb_renderer* r : m->r) { int rc = rtb_reset_counters(r); if (rc) return rc; }
	m->reduce_launches = 0;
	return RTB_OK;
}

}  // extern "C"

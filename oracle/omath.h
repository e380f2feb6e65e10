// oracle/omath.h — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/.
//
// Independent CPU restatement of the arithmetic spec in DESIGN.md ("Arithmetic spec"): binary32
// +,-,*,/,sqrt and explicit fmaf in a fixed order, so that the oracle and the CUDA kernels
// (compiled with -fmad=false) agree bit for bit on geometry.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct V3 {
	float x, y, z;
	V3() : x(0), y(0), z(0) {}
	V3(float a, float b, float c) : x(a), y(b), z(c) {}
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }

// dot: x*x' first, then fma in y, then fma in z  (glm::dot contracted)
inline float dot(V3 a, V3 b) { return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) {
	return V3(std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x)));
}
// a*s + b per component, fused  (Ray::at  ray_data.cuh:14)
inline V3 fma3(V3 a, float s, V3 b) { return V3(std::fmaf(a.x, s, b.x), std::fmaf(a.y, s, b.y), std::fmaf(a.z, s, b.z)); }
// glm::normalize = v * (1/sqrt(dot(v,v)))   Libraries/include/glm/detail/func_geometric.inl:81-90
inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
// glm::mix(x,y,a) = x*(1-a) + y*a
inline float mixf(float x, float y, float a) { return std::fmaf(y, a, x * (1.0f - a)); }
inline V3 mix(V3 x, V3 y, float a) { float b = 1.0f - a; return V3(std::fmaf(y.x, a, x.x * b), std::fmaf(y.y, a, x.y * b), std::fmaf(y.z, a, x.z * b)); }
// glm::reflect = I - N*dot(N,I)*2   func_geometric.inl:104-111
inline V3 reflect(V3 I, V3 N) { float k = dot(N, I) * 2.0f; return V3(std::fmaf(-N.x, k, I.x), std::fmaf(-N.y, k, I.y), std::fmaf(-N.z, k, I.z)); }
// glm::near_zero  main/src/utilities/glm_utils.h:15-25
inline bool near_zero(V3 a, float eps = 1e-9f) { return !(std::fabs(a.x) > eps) && !(std::fabs(a.y) > eps) && !(std::fabs(a.z) > eps); }
inline float length2(V3 a) { float s = 0.0f; s += a.x * a.x; s += a.y * a.y; s += a.z * a.z; return s; }  // glm_utils.h:27-35

// ---------------------------------------------------------------- Philox4x32-10 (Random123 / curand_philox4x32_x.h:88-91,160-185)
struct U4 { uint32_t v[4]; };
inline U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
	for (int round = 0; round < 10; ++round) {
		uint64_t p0 = (uint64_t)0xD2511F53u * ctr.v[0];
		uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr.v[2];
		U4 nx;
		nx.v[0] = (uint32_t)(p1 >> 32) ^ ctr.v[1] ^ k0;
		nx.v[1] = (uint32_t)p1;
		nx.v[2] = (uint32_t)(p0 >> 32) ^ ctr.v[3] ^ k1;
		nx.v[3] = (uint32_t)p0;
		ctr = nx;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	return ctr;
}
// curand_uniform mapping  /usr/local/cuda/include/curand_uniform.h:69-72 : (0,1]
inline float uniform01(uint32_t x) { return std::fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

struct F4 { float x, y, z, w; };
enum { STREAM_SCATTER = 0, STREAM_LENS = 1, STREAM_MEDIUM0 = 16 };
const uint32_t CAMERA_BOUNCE = 0xFFFFFFFFu;
inline F4 rng4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t stream) {
	U4 c; c.v[0] = sample; c.v[1] = bounce; c.v[2] = stream; c.v[3] = 0;
	U4 r = philox4x32_10(c, seed, pixel);
	return F4{uniform01(r.v[0]), uniform01(r.v[1]), uniform01(r.v[2]), uniform01(r.v[3])};
}

// ---------------------------------------------------------------- XORWOW (curand_kernel.h:800-826 init with subsequence 0, offset 0; :863-874 step)
struct Xorwow {
	uint32_t d, v[5];
	explicit Xorwow(uint64_t seed = 0) {
		uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
		uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
		d = 6615241u + t1 + t0;
		v[0] = 123456789u + t0; v[1] = 362436069u ^ t0; v[2] = 521288629u + t1; v[3] = 88675123u ^ t1; v[4] = 5783321u + t0;
	}
	uint32_t next_u32() {
		uint32_t t = v[0] ^ (v[0] >> 2);
		v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
		v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
		d += 362437u;
		return v[4] + d;
	}
	float next() { return uniform01(next_u32()); }   // cuRandom::next  cuRandom.cuh:21
};

// ---------------------------------------------------------------- sin/cos(2 pi u) for u in (0,1]: octant fold + Taylor kernels on [0, pi/4]
inline void sincos2pi(float u, float& sn, float& cs) {
	float x = u * 8.0f;
	int k = (int)x;
	float f = x - (float)k;
	int q = k & 7, odd = q & 1;
	float g = odd ? (1.0f - f) : f;
	float a = g * 0.78539816339744831f, a2 = a * a;
	float ps = std::fmaf(a2, 2.7557319223985893e-6f, -1.9841269841269841e-4f);
	ps = std::fmaf(a2, ps, 8.3333333333333332e-3f);
	ps = std::fmaf(a2, ps, -1.6666666666666666e-1f);
	float s = std::fmaf(a * a2, ps, a);
	float pc = std::fmaf(a2, -2.7557319223985888e-7f, 2.4801587301587302e-5f);
	pc = std::fmaf(a2, pc, -1.3888888888888889e-3f);
	pc = std::fmaf(a2, pc, 4.1666666666666664e-2f);
	pc = std::fmaf(a2, pc, -0.5f);
	float c = std::fmaf(a2, pc, 1.0f);
	int m = ((q + odd) >> 1) & 3;
	float sa = odd ? -s : s;
	if (m == 0) { sn = sa; cs = c; }
	else if (m == 1) { sn = c; cs = -sa; }
	else if (m == 2) { sn = -sa; cs = -c; }
	else { sn = -c; cs = sa; }
}
inline float sin_any(float x) {
	float r = x * 0.15915494309189535f;
	r = r - std::floor(r);
	float s, c; sincos2pi(r, s, c); return s;
}
// natural log of a positive normal float: exponent split + atanh series
inline float logpos(float x) {
	uint32_t b; std::memcpy(&b, &x, 4);
	int e = (int)(b >> 23) - 127;
	uint32_t mb = (b & 0x007FFFFFu) | 0x3F800000u;
	float m; std::memcpy(&m, &mb, 4);
	if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
	float s = (m - 1.0f) / (m + 1.0f), s2 = s * s;
	float p = std::fmaf(s2, 0.1111111111f, 0.1428571429f);
	p = std::fmaf(s2, p, 0.2f);
	p = std::fmaf(s2, p, 0.3333333333f);
	p = std::fmaf(s2, p, 1.0f);
	p = (2.0f * s) * p;
	return std::fmaf((float)e, 0.69314718056f, p);
}
inline V3 unit_sphere(float u0, float u1) {
	float z = std::fmaf(-2.0f, u0, 1.0f);
	float r2 = std::fmaf(-z, z, 1.0f);
	float r = std::sqrt(r2 < 0.0f ? 0.0f : r2);
	float s, c; sincos2pi(u1, s, c);
	return V3(r * c, r * s, z);
}
inline void unit_disc(float u0, float u1, float& dx, float& dy) {
	float r = std::sqrt(u0); float s, c; sincos2pi(u1, s, c); dx = r * c; dy = r * s;
}

}  // namespace orc

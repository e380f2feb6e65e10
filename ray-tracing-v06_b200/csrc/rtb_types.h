// rtb_types.h — flattened, aligned device records produced by the host flattener
// (rtb_scene.cpp) and consumed by the wavefront kernels (rtb_kernels.cu).
#ifndef RTB_TYPES_H
#define RTB_TYPES_H

#include <stdint.h>
#include <vector>

namespace rtb {

// Primitive types (3 bits of a leaf reference).
enum PrimType : int {
	PRIM_SPHERE = 0, PRIM_MOVING_SPHERE = 1, PRIM_QUAD = 2, PRIM_TRIANGLE = 3,
	PRIM_MEDIUM_SPHERE = 4, PRIM_MEDIUM_BOX = 5
};

// Child reference of a wide-layout BVH node: >= 0 inner node index; < 0 leaf,
// ~ref = (prim_index << 3) | PrimType.
inline int make_leaf_ref(int prim, int type) { return ~((prim << 3) | type); }

// 64-byte inner node holding BOTH children's boxes, fetched as 4 x LDG.128:
//   a = (l.min.x, l.min.y, l.min.z, l.max.x)   b = (l.max.y, l.max.z, r.min.x, r.min.y)
//   c = (r.min.z, r.max.x, r.max.y, r.max.z)   d = (left ref, right ref, 0, 0) as int bits
struct alignas(64) DevNode { float f[12]; int32_t left, right, pad0, pad1; };

// 64-byte primitive record (4 x float4):
//   SPHERE         q0 = (c.xyz, r)        q1 = (cos, sin, 0, 0) of the baked rotate_y (uv frame)
//   MOVING_SPHERE  q0 = (c0.xyz, r)       q1 = (c1.xyz, 0)   q2 = (cos, sin, 0, 0)
//   QUAD/TRIANGLE  q0 = (Q.xyz, D)        q1 = (u.xyz, N.x)  q2 = (v.xyz, N.y)  q3 = (w.xyz, N.z)
//   MEDIUM_SPHERE  q0 = (c.xyz, r)        q1 = (-1/density, medium index bits, 0, 0)
//   MEDIUM_BOX     q0 = (bmin.xyz, cos)   q1 = (bmax.xyz, sin)  q2 = (offset.xyz, -1/density)
//                  q3 = (medium index bits, 0, 0, 0)       world = R_y * object + offset
struct alignas(64) DevPrim { float q[16]; };

struct alignas(8) DevPrimInfo { int32_t material; int32_t object; };

struct alignas(16) DevMaterial { int32_t kind, tex; float param, pad; float albedo[4]; };  // 32 B

struct alignas(16) DevTexture {                                                             // 48 B
	int32_t kind, even, odd; float scale;      // checker: scale holds inv_scale
	float rgb[4];
	int32_t width, height; uint32_t blob_offset; uint32_t pad;
};

// Everything the kernels need about a scene, host-side staging.
struct FlatScene {
	std::vector<DevNode> nodes;          // wide layout actually traversed
	int32_t root_ref = 0;                // inner index or leaf ref
	std::vector<DevPrim> prims;
	std::vector<DevPrimInfo> prim_info;
	std::vector<int32_t> prim_type;
	std::vector<DevMaterial> materials;
	std::vector<DevTexture> textures;
	std::vector<uint8_t> blob;
	int32_t n_media = 0;
	int32_t background_mode = 0;
	float background[3] = {0, 0, 0};
	int32_t max_depth_nodes = 0;         // tree depth (stack bound check)
};

}  // namespace rtb
#endif

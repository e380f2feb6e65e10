/*
 * rtb_scene_format.h — the flat binary scene description ("RTBS" v1) written by
 * rtb_scene_serialize().  It is the only thing the product library and the CPU oracle
 * share: the oracle parses this blob and evaluates the *object graph* the way the
 * reference / the book does (recursive lists, BVHs, ray-transforming instances, two
 * boundary queries per medium), while the product flattens the same graph into
 * world-space SoA buffers.  All records are little-endian, 4-byte aligned.
 *
 *   rtbs_header
 *   rtbs_texture  [n_textures]
 *   rtbs_material [n_materials]
 *   rtbs_object   [n_objects]
 *   int32         children[n_children]     (LIST / BVH members)
 *   uint8         blob[n_blob_bytes]       (image texels RGB8; Perlin tables)
 */
#ifndef RTB_SCENE_FORMAT_H
#define RTB_SCENE_FORMAT_H

#include <stdint.h>

#define RTBS_MAGIC 0x53425452u /* "RTBS" */
#define RTBS_VERSION 1u

typedef struct rtbs_header {
	uint32_t magic, version;
	uint32_t n_textures, n_materials, n_objects, n_children;
	uint32_t n_blob_bytes;
	int32_t  root_object;
	int32_t  background_mode;   /* rtb_background_mode */
	float    background[3];
} rtbs_header;

/* Perlin tables in the blob: 256 x float[3] gradients, then perm_x/y/z: 3 x 256 x int32. */
#define RTBS_PERLIN_POINTS 256
#define RTBS_PERLIN_BYTES (256 * 3 * 4 + 3 * 256 * 4)

typedef struct rtbs_texture {
	int32_t  kind;          /* rtb_texture_kind */
	int32_t  even, odd;     /* checker children (texture ids) */
	float    scale;         /* checker: scale (inv_scale = 1/scale); noise: frequency */
	float    rgb[3];        /* solid colour */
	int32_t  width, height; /* image */
	uint32_t blob_offset;   /* image texels / Perlin tables */
	uint32_t seed;          /* noise */
	uint32_t pad;
} rtbs_texture;             /* 48 B */

typedef struct rtbs_material {
	int32_t kind;           /* rtb_material_kind */
	int32_t tex;            /* albedo / emission / phase texture id, -1 = use `albedo` */
	float   albedo[3];
	float   param;          /* metal: fuzz; dielectric: ior */
} rtbs_material;            /* 24 B */

/* f[] per kind:
 *   SPHERE           center.xyz, radius
 *   MOVING_SPHERE    center0.xyz, radius, center1.xyz
 *   QUAD / TRIANGLE  Q.xyz, u.xyz, v.xyz
 *   BOX              a.xyz, b.xyz          (six quads, outward normals)
 *   TRANSLATE        offset.xyz            child = children[child_begin]
 *   ROTATE_Y         degrees, sin, cos     child = children[child_begin]
 *   CONSTANT_MEDIUM  density, -1/density   boundary = children[child_begin]; mat = phase material
 *   LIST / BVH       (none)                members = children[child_begin .. +child_count); BVH: builder in `aux`
 */
typedef struct rtbs_object {
	int32_t kind;           /* rtb_object_kind */
	int32_t mat;
	int32_t child_begin, child_count;
	int32_t aux;
	float   f[11];
} rtbs_object;              /* 64 B */

#endif

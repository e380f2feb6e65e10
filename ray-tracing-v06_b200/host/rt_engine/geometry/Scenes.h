// Host mirror of main/src/rt_engine/geometry/Scenes.h:55-92 (SceneBook2BVH and its Factory, same
// public surface) plus a registry of the scenes BASELINE.json's configs name.  Definitions in
// host/Scenes.cpp.
#pragma once
#include <vector>

#include "rtb.h"

class aabb;
class Hittable;
class cuHostRND;
class BVH_Handle;
class SphereHandle;

class SceneBook2BVH {
	SceneBook2BVH();
	void _delete();
	BVH_Handle* bvh;
	aabb* world_bounds;
	std::vector<SphereHandle> sphere_handles;

public:
	~SceneBook2BVH();
	SceneBook2BVH(SceneBook2BVH&& scene);
	SceneBook2BVH& operator=(SceneBook2BVH&& scene);
	class Factory;
	const Hittable* getWorldPtr() const;
};

class SceneBook2BVH::Factory {
	void _delete();
	void _populate_world();
	cuHostRND* host_rnd;
	std::vector<SphereHandle> sphere_handles;

public:
	Factory();
	~Factory();
	Factory(Factory&& factory);
	Factory& operator=(Factory&& factory);
	SceneBook2BVH* MakeScene();
};

// ---- C accessors used by the Python tests / bench (exported from librtb200_scenes.so)
extern "C" {
typedef struct rtb_scene_info {
	rtb_camera camera;
	int32_t width, height, spp, max_depth;
} rtb_scene_info;

int rtb_scenes_count(void);
const char* rtb_scenes_name(int index);
// Builds the named scene through the host mirror API; the caller owns the returned rtb_scene
// (root and background already set).  NULL on error (rtb_last_error / rtb_scenes_last_error).
rtb_scene* rtb_scenes_build(const char* name, rtb_scene_info* info);
const char* rtb_scenes_last_error(void);
}

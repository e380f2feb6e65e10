// Host mirror of main/src/Renderer.{h,cu}: Renderer::MakeRenderer / Render / DownloadRenderbuffer
// with the reference's signatures (Renderer.h:38-46).  MakeRenderer flattens and uploads the scene
// that `d_world_ptr` is the root of; Render launches the wavefront kernels through rtb_render and
// waits (the reference's Render is synchronous too, Renderer.cu:132-133); the camera is read at
// Render time (Renderer.cu:117).  Overloads for the other two camera PODs are an extension.
#pragma once
#include <cstdint>
#include <cstdio>
#include <glm/glm.hpp>
#include <stdexcept>
#include <string>

#include "rt_engine/geometry/hittable.cuh"
#include "rt_engine/shaders/cu_Cameras.cuh"
#include "rtb_context.h"

class Renderer {
	struct M {
		uint32_t render_width{}, render_height{};
		uint32_t samples_per_pixel{}, max_depth{};
		const MotionBlurCamera* cam{};
		const DefocusBlurCamera* defocus_cam{};
		const PinholeCamera* pinhole_cam{};
		rtb_renderer* r{};
		rtb_scene* scene{};
		uint32_t seed{1984};
	} m;

	void _delete() { if (m.r) rtb_renderer_destroy(m.r); m.r = nullptr; }
	Renderer(M m) : m(m) {}
	Renderer(const Renderer&) = delete;
	Renderer& operator=(const Renderer&) = delete;

	static Renderer _make(uint32_t w, uint32_t h, uint32_t spp, uint32_t depth, const Hittable* world, M m) {
		rtb_scene* s = rtb_host::scene();
		rtb_host::check(rtb_scene_set_root(s, world->rtb_object), "rtb_scene_set_root");
		rtb_host::check(rtb_renderer_create(&m.r, 0), "rtb_renderer_create");
		rtb_host::check(rtb_renderer_set_scene(m.r, s), "rtb_renderer_set_scene");
		m.render_width = w; m.render_height = h; m.samples_per_pixel = spp; m.max_depth = depth; m.scene = s;
		return Renderer(m);
	}

public:
	~Renderer() { _delete(); }
	Renderer(Renderer&& o) noexcept : m(o.m) { o.m.r = nullptr; }
	Renderer& operator=(Renderer&& o) noexcept { if (this != &o) { _delete(); m = o.m; o.m.r = nullptr; } return *this; }

	static Renderer MakeRenderer(uint32_t render_width, uint32_t render_height, uint32_t samples_per_pixel, uint32_t max_depth,
	                             const MotionBlurCamera* cam, const Hittable* d_world_ptr) {
		M m{}; m.cam = cam; return _make(render_width, render_height, samples_per_pixel, max_depth, d_world_ptr, m);
	}
	static Renderer MakeRenderer(uint32_t render_width, uint32_t render_height, uint32_t samples_per_pixel, uint32_t max_depth,
	                             const DefocusBlurCamera* cam, const Hittable* d_world_ptr) {
		M m{}; m.defocus_cam = cam; return _make(render_width, render_height, samples_per_pixel, max_depth, d_world_ptr, m);
	}
	static Renderer MakeRenderer(uint32_t render_width, uint32_t render_height, uint32_t samples_per_pixel, uint32_t max_depth,
	                             const PinholeCamera* cam, const Hittable* d_world_ptr) {
		M m{}; m.pinhole_cam = cam; return _make(render_width, render_height, samples_per_pixel, max_depth, d_world_ptr, m);
	}

	void Render() {
		rtb_camera c = m.cam ? m.cam->to_rtb() : (m.defocus_cam ? m.defocus_cam->to_rtb() : m.pinhole_cam->to_rtb());
		rtb_host::check(rtb_renderer_set_camera(m.r, &c), "rtb_renderer_set_camera");
		rtb_render_params p{};
		p.width = m.render_width; p.height = m.render_height; p.sample_begin = 0; p.sample_end = m.samples_per_pixel;
		p.max_depth = m.max_depth; p.seed = m.seed; p.flags = RTB_RENDER_CLEAR;
		printf("Running render kernel...\n");
		rtb_host::check(rtb_render(m.r, &p, nullptr), "rtb_render");
		rtb_host::check(rtb_synchronize(m.r), "rtb_synchronize");
		rtb_counters k{}; rtb_get_counters(m.r, &k);
		printf("Rendering finished in %fms.\n", k.render_ms);
	}
	void DownloadRenderbuffer(glm::vec4* host_dst) const { rtb_host::check(rtb_download(m.r, &host_dst->x), "rtb_download"); }
	rtb_renderer* handle() const { return m.r; }
};

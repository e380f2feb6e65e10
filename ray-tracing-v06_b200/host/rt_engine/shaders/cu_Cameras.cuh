// Host mirror of main/src/rt_engine/shaders/cu_Cameras.cuh:12-90: the three camera PODs with the
// reference's constructor signatures.  Basis construction is delegated to the C ABI
// (rtb_camera_* in csrc/rtb_scene.cpp); sample_ray runs in the generate kernel.
#pragma once
#include <glm/glm.hpp>

#include "rtb.h"

namespace rtb_host {
inline glm::vec3 v3_of(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
}

struct PinholeCamera {
	glm::vec3 o{}, u{}, v{}, w{};
	PinholeCamera() {}
	PinholeCamera(glm::vec3 lookfrom, glm::vec3 lookat, glm::vec3 up, float vfov, float aspect_ratio) {
		rtb_camera c; rtb_camera_pinhole(&c, &lookfrom.x, &lookat.x, &up.x, vfov, aspect_ratio);
		o = rtb_host::v3_of(c.o); u = rtb_host::v3_of(c.u); v = rtb_host::v3_of(c.v); w = rtb_host::v3_of(c.w);
	}
	Ray sample_ray(float s, float t) const { return Ray(o, w + u * s + v * t); }
	rtb_camera to_rtb() const {
		rtb_camera c{}; c.kind = RTB_CAM_PINHOLE;
		for (int i = 0; i < 3; ++i) { c.o[i] = o[i]; c.u[i] = u[i]; c.v[i] = v[i]; c.w[i] = w[i]; }
		return c;
	}
};

struct DefocusBlurCamera {
	glm::vec3 o{}, u{}, v{}, w{};
	float viewport_width{}, viewport_height{};
	float lens_radius{}, focus_dist{};
	DefocusBlurCamera() {}
	DefocusBlurCamera(glm::vec3 lookfrom, glm::vec3 lookat, glm::vec3 up, float vfov, float aspect_ratio, float aperture, float focus_dist) {
		rtb_camera c; rtb_camera_defocus(&c, &lookfrom.x, &lookat.x, &up.x, vfov, aspect_ratio, aperture, focus_dist, 0.0f, 0.0f);
		o = rtb_host::v3_of(c.o); u = rtb_host::v3_of(c.u); v = rtb_host::v3_of(c.v); w = rtb_host::v3_of(c.w);
		viewport_width = c.viewport_width; viewport_height = c.viewport_height; lens_radius = c.lens_radius; this->focus_dist = c.focus_dist;
	}
	rtb_camera to_rtb() const {
		rtb_camera c{}; c.kind = RTB_CAM_DEFOCUS;
		for (int i = 0; i < 3; ++i) { c.o[i] = o[i]; c.u[i] = u[i]; c.v[i] = v[i]; c.w[i] = w[i]; }
		c.viewport_width = viewport_width; c.viewport_height = viewport_height; c.lens_radius = lens_radius; c.focus_dist = focus_dist;
		return c;
	}
};

struct MotionBlurCamera {
	glm::vec3 o, u, v, w;
	float t0, t1;
	MotionBlurCamera() : o(), u(), v(), w(), t0(0.0f), t1(1.0f) {}
	MotionBlurCamera(glm::vec3 lookfrom, glm::vec3 lookat, glm::vec3 up, float vfov, float aspect_ratio, float time0, float time1) {
		rtb_camera c; rtb_camera_motion(&c, &lookfrom.x, &lookat.x, &up.x, vfov, aspect_ratio, time0, time1);
		o = rtb_host::v3_of(c.o); u = rtb_host::v3_of(c.u); v = rtb_host::v3_of(c.v); w = rtb_host::v3_of(c.w);
		t0 = time0; t1 = time1;
	}
	rtb_camera to_rtb() const {
		rtb_camera c{}; c.kind = RTB_CAM_MOTION;
		for (int i = 0; i < 3; ++i) { c.o[i] = o[i]; c.u[i] = u[i]; c.v[i] = v[i]; c.w[i] = w[i]; }
		c.t0 = t0; c.t1 = t1; c.focus_dist = 1.0f;
		return c;
	}
};

// rtb_app — a Linux command-line shell around the renderer, written against the host mirror the way
// the reference's RT application is (main/src/main.cpp:22-32, main/src/FirstApp.cpp:20-56,94-122):
// build the camera, build the scene, MakeRenderer, Render, DownloadRenderbuffer, write the image.
//
//   rtb_app [scene] [--width W] [--height H] [--spp N] [--depth D] [--out image.ppm|image.pfm]
//   rtb_app --obj model.obj [...]     a Wavefront OBJ mesh (MeshHandle) on a checkered floor under the sky
//
// scene = one of rtb_scenes_name(i) (default book2_bouncing, which goes through SceneBook2BVH::Factory
// and Renderer::MakeRenderer exactly like FirstApp::MakeApp).  The 8-bit writer follows
// write_renderbuffer (FirstApp.cpp:108-122): uint8 = value * 255.999f, RGB, rows flipped (row 0 of the
// float buffer is the bottom of the image); .pfm keeps the raw floats.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Renderer.h"
#include "rt_engine/geometry/BVH.cuh"
#include "rt_engine/geometry/HittableList.cuh"
#include "rt_engine/geometry/Mesh.cuh"
#include "rt_engine/geometry/Scenes.h"
#include "rt_engine/geometry/SphereHittable.cuh"
#include "rt_engine/shaders/cu_Cameras.cuh"
#include "rt_engine/shaders/cu_Textures.cuh"
#include "rt_engine/shaders/cu_materials.cuh"

static bool write_image(const std::string& path, uint32_t width, uint32_t height, const std::vector<glm::vec4>& data) {
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	const bool pfm = path.size() > 4 && path.substr(path.size() - 4) == ".pfm";
	if (pfm) {   // PFM stores rows bottom-up, like the render buffer
		fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
		for (uint32_t i = 0; i < width * height; ++i) fwrite(&data[i].x, sizeof(float), 3, f);
	} else {
		fprintf(f, "P6\n%u %u\n255\n", width, height);
		for (uint32_t y = 0; y < height; ++y) {
			const glm::vec4* row = &data[(size_t)(height - 1 - y) * width];   // stbi_flip_vertically_on_write(true)
			for (uint32_t x = 0; x < width; ++x) {
				unsigned char px[3] = {static_cast<unsigned char>(row[x][0] * 255.999f), static_cast<unsigned char>(row[x][1] * 255.999f),
				                       static_cast<unsigned char>(row[x][2] * 255.999f)};
				fwrite(px, 1, 3, f);
			}
		}
	}
	fclose(f);
	return true;
}

int main(int argc, char** argv) {
	std::string scene_name = "book2_bouncing", out = "render.ppm", obj_path;
	int width = 0, height = 0, spp = 0, depth = 0;
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
		if (a == "--width") width = next();
		else if (a == "--height") height = next();
		else if (a == "--spp") spp = next();
		else if (a == "--depth") depth = next();
		else if (a == "--out" && i + 1 < argc) out = argv[++i];
		else if (a == "--obj" && i + 1 < argc) obj_path = argv[++i];
		else if (a == "--list") { for (int k = 0; k < rtb_scenes_count(); ++k) printf("%s\n", rtb_scenes_name(k)); return 0; }
		else scene_name = a;
	}
	try {
		std::vector<glm::vec4> fb;
		if (!obj_path.empty()) {
			// A mesh scene assembled the way FirstApp assembles its world: materials, handles, a HittableList, MakeRenderer.
			uint32_t _width = width ? width : 800, _height = height ? height : 600;
			printf("Loading %s... ", obj_path.c_str());
			Mesh mesh = MeshHandle::LoadObj(obj_path);
			Lambertian clay(glm::vec3(0.73f, 0.73f, 0.73f));
			MeshHandle model = MeshHandle::MakeMesh(mesh, &clay);
			printf("%d triangles.\n", model.triangleCount());
			const glm::vec3 lo = model.getBounds().getMin(), hi = model.getBounds().getMax(), mid = (lo + hi) * 0.5f;
			const float size = glm::length(hi - lo);
			solid_texture dark(glm::vec3(.2f, .3f, .1f)), light(glm::vec3(.9f, .9f, .9f));
			checker_texture checks(&dark, &light, 8.0f / size);
			Lambertian floor_mat(&checks);
			GeoHandle floor = GeoHandle::MakeQuad(Quad(glm::vec3(mid.x - 10 * size, lo.y, mid.z - 10 * size), glm::vec3(20 * size, 0, 0), glm::vec3(0, 0, 20 * size)), &floor_mat);
			const Hittable* objs[2] = {floor.getHittablePtr(), model.getHittablePtr()};
			aabb bounds = model.getBounds(); bounds += floor.getBounds();
			HittableList world(objs, 2, bounds);
			MotionBlurCamera cam(mid + glm::vec3(0.55f, 0.35f, 1.1f) * size, mid, glm::vec3(0, 1, 0), 40.0f, _width / (float)_height, 0.0f, 1.0f);
			Renderer renderer = Renderer::MakeRenderer(_width, _height, spp ? spp : 64, depth ? depth : 16, &cam, &world);
			renderer.Render();
			fb.resize((size_t)_width * _height);
			renderer.DownloadRenderbuffer(fb.data());
			width = _width; height = _height;
		} else if (scene_name == "book2_bouncing") {
			// FirstApp::MakeApp, statement for statement, with the resolution / spp / depth made arguments
			uint32_t _width = width ? width : 1280, _height = height ? height : 720;
			printf("Building MotionBlurCamera object... ");
			auto cam = std::make_unique<MotionBlurCamera>(glm::vec3(13, 2, 3), glm::vec3(0, 0, 0), glm::vec3(0, 1, 0), 30.0f, _width / (float)_height, 0.1f, 1.0f);
			printf("done.\nBuilding SceneBook2BVH object...\n");
			SceneBook2BVH::Factory scene_factory{};
			std::unique_ptr<SceneBook2BVH> scene_ptr(scene_factory.MakeScene());
			printf("SceneBook2BVH object built.\nMaking Renderer object...\n");
			Renderer renderer = Renderer::MakeRenderer(_width, _height, spp ? spp : 1, depth ? depth : 4, cam.get(), scene_ptr->getWorldPtr());
			printf("Renderer object built.\nRendering scene...\n");
			renderer.Render();
			fb.resize((size_t)_width * _height);
			renderer.DownloadRenderbuffer(fb.data());
			width = _width; height = _height;
		} else {
			rtb_scene_info info{};
			rtb_scene* s = rtb_scenes_build(scene_name.c_str(), &info);
			if (!s) { fprintf(stderr, "rtb_app: %s\n", rtb_scenes_last_error()); return 1; }
			if (!width) width = info.width;
			if (!height) height = info.height;
			rtb_renderer* r = nullptr;
			rtb_host::check(rtb_renderer_create(&r, 0), "rtb_renderer_create");
			rtb_host::check(rtb_renderer_set_scene(r, s), "rtb_renderer_set_scene");
			rtb_host::check(rtb_renderer_set_camera(r, &info.camera), "rtb_renderer_set_camera");
			rtb_render_params p{};
			p.width = width; p.height = height; p.sample_end = spp ? spp : info.spp; p.max_depth = depth ? depth : info.max_depth; p.seed = 1984; p.flags = RTB_RENDER_CLEAR;
			printf("Running render kernel...\n");
			rtb_host::check(rtb_render(r, &p, nullptr), "rtb_render");
			rtb_host::check(rtb_synchronize(r), "rtb_synchronize");
			rtb_counters k{}; rtb_get_counters(r, &k);
			printf("Rendering finished in %fms (%.1f Mrays/s, %.1f Mpaths/s).\n", k.render_ms, k.rays / k.render_ms * 1e-3, k.paths / k.render_ms * 1e-3);
			fb.resize((size_t)width * height);
			rtb_host::check(rtb_download(r, &fb[0].x), "rtb_download");
			rtb_renderer_destroy(r); rtb_scene_destroy(s);
		}
		printf("Writing render to disk... ");
		if (!write_image(out, width, height, fb)) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
		printf("done (%s).\n", out.c_str());
	} catch (const std::exception& e) {
		fprintf(stderr, "rtb_app: %s\n", e.what());
		return 1;
	}
	return 0;
}

// Host mirror of main/src/rt_engine/shaders/cu_Textures.cuh:9-40 plus the Book 2 textures the
// reference has not reached (image_texture, noise_texture).
#pragma once
#include <cstdint>
#include <glm/glm.hpp>

#include "../../rtb_context.h"
#include "texture.cuh"

class solid_texture : public Texture {
public:
	solid_texture() = default;
	solid_texture(glm::vec3 color) { rtb_texture = rtb_host::check(rtb_add_solid_texture(rtb_host::scene(), &color.x), "solid_texture"); }
};

// checker_texture::value keeps the reference's truncation + C '%' parity rule (cu_Textures.cuh:32-39).
class checker_texture : public Texture {
public:
	checker_texture() = default;
	checker_texture(solid_texture* c1, solid_texture* c2, float scale) : checker_texture(static_cast<Texture*>(c1), static_cast<Texture*>(c2), scale) {}
	checker_texture(Texture* even, Texture* odd, float scale) {
		rtb_texture = rtb_host::check(rtb_add_checker_texture(rtb_host::scene(), scale, even->rtb_texture, odd->rtb_texture), "checker_texture");
	}
};

class image_texture : public Texture {
public:
	image_texture(const uint8_t* pixels, int width, int height, int channels) {
		rtb_texture = rtb_host::check(rtb_add_image_texture(rtb_host::scene(), pixels, width, height, channels), "image_texture");
	}
};

class noise_texture : public Texture {
public:
	noise_texture(float scale, uint32_t seed = 1984) { rtb_texture = rtb_host::check(rtb_add_noise_texture(rtb_host::scene(), scale, seed), "noise_texture"); }
};

// Host mirror of main/src/rt_engine/shaders/texture.cuh:8-20.
#pragma once
class Texture {
protected:
	Texture() = default;

public:
	virtual ~Texture() = default;
	int rtb_texture = -1;
};

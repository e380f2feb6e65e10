// Host mirror of main/src/utilities/cuda_utilities/cuHostRND.{h,cpp}: a buffer of XORWOW uniforms
// (seed, offset 0, default ordering) that doubles when exhausted (cuHostRND.cpp:57-65).  Uses the
// cuRAND *host* generator, which yields the same sequence as the reference's device generator and
// needs no GPU.
#pragma once
#include <curand.h>

#include <cstddef>
#include <stdexcept>
#include <vector>

class cuHostRND {
	enum { MULTIPLIER_FACTOR = 2 };
	std::vector<float> rnd_uniforms;
	size_t head = 0;
	curandGenerator_t gen = nullptr;

	void _populate_buffer() {
		if (curandGenerateUniform(gen, rnd_uniforms.data(), rnd_uniforms.size()) != CURAND_STATUS_SUCCESS)
			throw std::runtime_error("curandGenerateUniform failed");
	}

public:
	cuHostRND(const cuHostRND&) = delete;
	cuHostRND& operator=(const cuHostRND&) = delete;

	cuHostRND(size_t capacity, size_t seed, size_t offset = 0, curandOrdering_t ordering = CURAND_ORDERING_PSEUDO_DEFAULT)
	    : rnd_uniforms(capacity) {
		if (curandCreateGeneratorHost(&gen, CURAND_RNG_PSEUDO_XORWOW) != CURAND_STATUS_SUCCESS) throw std::runtime_error("curandCreateGeneratorHost failed");
		curandSetPseudoRandomGeneratorSeed(gen, seed);
		curandSetGeneratorOffset(gen, offset);
		curandSetGeneratorOrdering(gen, ordering);
		_populate_buffer();
	}
	~cuHostRND() { if (gen) curandDestroyGenerator(gen); }

	float next() {
		if (head == rnd_uniforms.size()) {
			rnd_uniforms.resize(rnd_uniforms.size() * MULTIPLIER_FACTOR);
			head = 0;
			_populate_buffer();
		}
		return rnd_uniforms[head++];
	}
};

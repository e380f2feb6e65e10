// Host mirror of main/src/utilities/cuda_utilities/cuda_utils.cuh:16-23.
// newOnDevice<T>(args...) keeps its name and meaning ("construct a T the renderer can use and give
// me a pointer to it") but there is no per-object cudaMalloc + <<<1,1>>> launch + synchronize any
// more: T's constructor records a descriptor in the current rtb scene and the whole scene is
// uploaded once, as one arena, by Renderer::MakeRenderer.
#pragma once
#include <utility>

template <typename T, typename... Args>
inline T* newOnDevice(const Args&... args) { return new T(args...); }

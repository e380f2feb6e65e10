// Host mirror of main/src/utilities/cuda_utilities/cuError.h:22-29.  The reference's CUDA_ASSERT wraps raw
// CUDA runtime calls in scene code; behind the C ABI there are none left, so the macros only evaluate.
#pragma once
#define CUDA_ASSERT(expr) ((void)(expr))
#define CUDA_CHECK(expr) ((void)(expr))

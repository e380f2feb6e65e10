// rtb_context.h — the implicit scene the host mirror records into.
//
// In the reference, newOnDevice<T>() and the *Handle factories allocate device objects in a global
// address space (main/src/utilities/cuda_utilities/cuda_utils.cuh:16-23); there is no scene object
// to pass around.  The mirror keeps that calling convention: every constructor below records a
// descriptor into the thread's current rtb_scene through the C ABI (include/rtb.h).
#pragma once
#include <stdexcept>
#include <string>

#include "rtb.h"

namespace rtb_host {

inline rtb_scene*& current_scene_slot() { static thread_local rtb_scene* s = nullptr; return s; }

// The scene being assembled (created on first use).
inline rtb_scene* scene() {
	rtb_scene*& s = current_scene_slot();
	if (!s && rtb_scene_create(&s) != RTB_OK) throw std::runtime_error(std::string("rtb_scene_create: ") + rtb_last_error());
	return s;
}
// Detach the current scene (the caller now owns it); the next constructor starts a new one.
inline rtb_scene* release_scene() { rtb_scene* s = current_scene_slot(); current_scene_slot() = nullptr; return s; }
inline void new_scene() { rtb_scene* s = release_scene(); if (s) rtb_scene_destroy(s); }

inline int check(int rc, const char* what) {
	if (rc < 0) throw std::runtime_error(std::string(what) + ": " + rtb_last_error());
	return rc;
}

}  // namespace rtb_host

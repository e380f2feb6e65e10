// Host mirror of main/src/rt_engine/shaders/material.cuh:15-54.
//
// What scene code needs from this header is (1) a common `Material` base to hold in handles,
// (2) the two families the reference distinguishes — materials written against one geometry type
// (`GeometryDependantMaterial<G>`, e.g. LambertianAbstract<Sphere>) and geometry-agnostic ones
// (`GeoIndependantMaterial`) — and (3) the `GeoAcceptableMat<G, M>` constraint the *Handle factories
// are declared with.  On the host a material is only a registered descriptor: `rtb_material` is its
// id in the scene being assembled; Scatter lives in csrc/rtb_kernels.cu.
#pragma once
#include <concepts>

#include "../geometry/hittable.cuh"

class Material {
public:
	int rtb_material = -1;          // id returned by rtb_add_<material>()
	virtual ~Material() = default;

protected:                          // only concrete materials can be created, as in the reference
	Material() = default;
	Material(const Material&) = default;
	Material& operator=(const Material&) = default;
};

// usable with any geometry
class GeoIndependantMaterial : public Material {};

// written for geometry type G
template <Geometry_t G>
class GeometryDependantMaterial : public Material {};

template <typename MatType>
concept GeoIndependantMat = std::derived_from<MatType, GeoIndependantMaterial>;

template <typename GeoType, typename MatType>
concept GeoDependantMat = std::derived_from<MatType, GeometryDependantMaterial<GeoType>>;

// a material M may be attached to geometry G
template <typename GeoType, typename MatType>
concept GeoAcceptableMat = GeoIndependantMat<MatType> || GeoDependantMat<GeoType, MatType>;

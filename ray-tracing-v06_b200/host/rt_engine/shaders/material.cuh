// Host mirror of main/src/rt_engine/shaders/material.cuh:15-54: the material class lattice and the
// GeoAcceptableMat concept that SphereHandle::MakeSphere & co. are constrained with.
#pragma once
#include <concepts>

#include "../geometry/hittable.cuh"

class Material {
protected:
	Material() = default;
	Material(const Material&) = default;
	Material& operator=(const Material&) = default;

public:
	virtual ~Material() = default;
	int rtb_material = -1;
};

class GeoIndependantMaterial : public Material {};
template <Geometry_t G>
class GeometryDependantMaterial : public Material {};

template <typename GeoType, typename MatType>
concept GeoDependantMat = std::derived_from<MatType, GeometryDependantMaterial<GeoType>>;
template <typename MatType>
concept GeoIndependantMat = std::derived_from<MatType, GeoIndependantMaterial>;
template <typename GeoType, typename MatType>
concept GeoAcceptableMat = GeoIndependantMat<MatType> || GeoDependantMat<GeoType, MatType>;

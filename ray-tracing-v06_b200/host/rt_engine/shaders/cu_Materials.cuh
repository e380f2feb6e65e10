// The reference includes this header under two spellings (Scenes.cu:16 asks for cu_Materials.cuh).
#pragma once
#include "cu_materials.cuh"

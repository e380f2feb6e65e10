"""Drop-in proof on the host side: the REFERENCE'S OWN scene builder, main/src/rt_engine/geometry/Scenes.cu,
compiled unchanged (as plain C++, no nvcc) against the host mirror headers and linked to librtb200.so, builds a
scene byte-identical to the one host/Scenes.cpp builds.  Needs /root/reference (build container only); the
reference file is copied into a temp directory for the compile and never into the repository."""
import shutil
import subprocess
from pathlib import Path

import pytest

from conftest import ROOT

REF_SCENES = Path("/root/reference/main/src/rt_engine/geometry/Scenes.cu")

MAIN = r'''
#include <cstdio>
#include <vector>
#include "rt_engine/geometry/Scenes.h"
#include "rt_engine/geometry/hittable.cuh"
#include "rtb_context.h"
int main(int argc, char** argv) {
	SceneBook2BVH::Factory factory{};                       // FirstApp.cpp:34-35
	SceneBook2BVH* scene = factory.MakeScene();
	rtb_scene* s = rtb_host::scene();
	rtb_scene_set_root(s, scene->getWorldPtr()->rtb_object);
	size_t n = rtb_scene_serialize(s, nullptr, 0);
	std::vector<unsigned char> buf(n);
	rtb_scene_serialize(s, buf.data(), n);
	FILE* o = fopen(argv[1], "wb"); fwrite(buf.data(), 1, n, o); fclose(o);
	return 0;
}
'''


@pytest.mark.skipif(not REF_SCENES.exists(), reason="needs the reference checkout")
def test_reference_scenes_cu_compiles_against_the_mirror(rtb, tmp_path):
    tree = tmp_path / "mirror"
    shutil.copytree(ROOT / "ray-tracing-v06_b200" / "host", tree)
    (tree / "rt_engine" / "geometry" / "Scenes_reference.cpp").write_bytes(REF_SCENES.read_bytes())   # unchanged
    (tree / "main.cpp").write_text(MAIN)
    lib_dir = ROOT / "ray-tracing-v06_b200"
    exe = tree / "dropin"
    cmd = ["g++", "-std=c++20", "-O1", "-I.", "-Iglm_compat", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include", "main.cpp",
           "rt_engine/geometry/Scenes_reference.cpp", f"-L{lib_dir}", "-lrtb200", "-L/usr/local/cuda/lib64", "-lcurand",
           f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", str(exe)]
    res = subprocess.run(cmd, cwd=tree, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    out = tmp_path / "scene.bin"
    subprocess.run([str(exe), str(out)], check=True, capture_output=True)
    assert out.read_bytes() == rtb.Scene.named("book2_bouncing").serialize()

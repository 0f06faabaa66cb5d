"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol that
include/b200rt.h declares, its host-side sutil mirrors agree with the reference golden vectors, and the
compute entry points fail loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import json
import pathlib
import re
import struct

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
KAT = json.loads((ROOT / "tests" / "golden" / "kat.json").read_text())


@pytest.fixture(scope="module")
def lib():
    from optix_raytracer_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from optix_raytracer_b200 import _lib
    header = (ROOT / "include" / "b200rt.h").read_text()
    declared = set(re.findall(r"^(?:int|void|uint64_t|const char\*)\s+(b200rt_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b200rt.h but not exported by libb200rt.so"
    # and the Python binding covers the whole header
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_struct_layouts_match_the_optix_sizes():
    from optix_raytracer_b200 import _lib, host
    assert C.sizeof(_lib.BuildInput) == 1032          # OptixBuildInput
    assert C.sizeof(_lib.TriangleArray) == 240        # OptixBuildInputTriangleArray
    assert _lib.TriangleArray.transformFormat.offset == 92
    assert C.sizeof(_lib.Instance) == 80              # OptixInstance
    assert C.sizeof(_lib.ShaderBindingTable) == 64    # OptixShaderBindingTable
    assert C.sizeof(_lib.AccelBuildOptions) == 20
    assert C.sizeof(host.PTParams) == 152 and host.PTParams.eye.offset == 36 and host.PTParams.light.offset == 84 and host.PTParams.handle.offset == 144
    assert C.sizeof(host.MGParams) == 168 and host.MGParams.eye.offset == 48 and host.MGParams.light.offset == 96 and host.MGParams.handle.offset == 160
    assert C.sizeof(host.RaycastParams) == 24


def test_error_strings(lib):
    assert lib.b200rt_error_name(0) == b"B200RT_SUCCESS"
    assert lib.b200rt_error_name(7001) == b"B200RT_ERROR_INVALID_VALUE"
    assert lib.b200rt_error_string(7900) == b"Error during CUDA call"
    assert b"sm_100a" in lib.b200rt_version()


def test_host_camera_matches_reference_golden():
    from optix_raytracer_b200 import host
    for e in KAT["camera_uvw"]:
        asp = struct.unpack("<f", struct.pack("<I", e["aspect_bits"]))[0]
        U, V, W = host.camera_uvw(e["eye"], e["lookat"], e["up"], e["fov_y"], asp)
        got = [struct.unpack("<I", struct.pack("<f", float(x)))[0] for x in list(U) + list(V) + list(W)]
        assert got == e["uvw_bits"]


def test_host_work_distribution_matches_reference_golden():
    from optix_raytracer_b200 import host
    for e in KAT["work_distribution"]:
        assert host.wd_num_samples(e["w"], e["h"], e["ngpu"]) == e["num_samples"]
        for s, x, y in e["pixels"]:
            assert host.wd_sample_pixel(e["w"], e["h"], e["ngpu"], e["gpu"], s) == (x, y)


def test_light_normal_matches_reference_golden():
    from optix_raytracer_b200 import host
    n = host.light_normal([0.0, 0.0, 105.0], [-130.0, 0.0, 0.0])
    assert [struct.unpack("<I", struct.pack("<f", float(x)))[0] for x in n] == KAT["cornell_light_normal_bits"]


def test_no_cpu_fallback(lib):
    """Without a CUDA device context creation fails with a CUDA error — nothing computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from optix_raytracer_b200 import _lib, host
    h = C.c_void_p()
    cb = _lib.LOG_CB(lambda *a: None)
    rc = lib.b200rt_context_create(0, cb, None, 0, C.byref(h))
    assert rc == 7900 and not h.value
    with pytest.raises(host.B200RTError):
        host.Context(0)
    # entry points reject a null context instead of computing anything
    assert lib.b200rt_trace_closest(None, None, 1, 1, 1, 0, 1) == 7051


def test_cornell_fixture_is_consistent():
    from optix_raytracer_b200 import host
    sc = host.load_cornell()
    assert sc["vertices"].shape == (96, 3) and sc["mat_indices"].shape == (32,)
    assert sc["mat_indices"].max() == 3 and list(sc["mat_indices"][-2:]) == [3, 3]
    # the emissive triangles are the ceiling light at y = 548.6
    light = sc["vertices"].reshape(32, 3, 3)[sc["mat_indices"] == 3]
    assert np.all(light[:, :, 1] == np.float32(548.6))

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
    declared = set(re.findall(r"^(?:int|void|uint64_t|unsigned int|const char\*)\s+(b200rt_[a-z0-9_]+)\s*\(", header, re.M))
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


def test_launch_struct_layouts_match_the_reference_headers():
    """Sizes / offsets measured on the reference's own headers (oracle/ref_shim.cpp -> tests/golden/kat.json) against the
    ctypes mirrors the host uses and the offsets csrc/raycast.cu reads whitted::HitGroupData with."""
    import ctypes as C
    import json
    import pathlib
    import re
    import struct as pystruct  # noqa: F401
    torch = pytest.importorskip("torch")  # host.py owns device memory through torch
    from optix_raytracer_b200 import host, _lib as L
    kat = json.loads((pathlib.Path(__file__).parent / "golden" / "kat.json").read_text())
    pl, hl = kat["params_layout"], kat["hitgroup_layout"]
    assert C.sizeof(host.PTParams) == pl["pt_Params"] and host.PTParams.eye.offset == pl["pt_Params_eye"]
    assert host.PTParams.light.offset == pl["pt_Params_light"] and host.PTParams.handle.offset == pl["pt_Params_handle"]
    assert C.sizeof(host.MGParams) == pl["mg_Params"] and host.MGParams.eye.offset == pl["mg_Params_eye"]
    assert host.MGParams.light.offset == pl["mg_Params_light"] and host.MGParams.handle.offset == pl["mg_Params_handle"]
    assert host.MGParams.sample_index_buffer.offset == pl["mg_Params_sample_index_buffer"] and host.MGParams.device_idx.offset == pl["mg_Params_device_idx"]
    assert C.sizeof(host.RaycastParams) == pl["rc_Params"] and pl["rc_Ray"] == 32 and pl["rc_Hit"] == 16
    assert (pl["pt_HitGroupData"], pl["pt_HitGroupData_diffuse_color"], pl["pt_HitGroupData_vertices"], pl["pt_MissData"]) == (32, 12, 24, 16)
    assert C.sizeof(L.BuildInput) == pl["OptixBuildInput"] and C.sizeof(L.Instance) == pl["OptixInstance"]
    assert C.sizeof(L.ShaderBindingTable) == pl["OptixShaderBindingTable"] and C.sizeof(L.AccelBuildOptions) == pl["OptixAccelBuildOptions"]
    assert C.sizeof(L.TriangleArray) == pl["OptixBuildInputTriangleArray"]
    # whitted::HitGroupData: the union of GeometryData is 16-byte aligned, so TriangleMesh starts at 16 (not 8)
    assert (host.HG_OFF_INDICES, host.HG_OFF_POSITIONS, host.HG_OFF_NORMALS, host.HG_OFF_MATERIAL, host.HG_SIZE) == \
        (hl["off_indices"], hl["off_positions"], hl["off_normals"], hl["off_material_data"], hl["sizeof_HitGroupData"])
    assert (hl["sizeof_BufferView"], hl["bv_off_data"], hl["bv_off_count"], hl["bv_off_byte_stride"], hl["bv_off_elmt_byte_size"]) == (16, 0, 8, 12, 14)
    src = (pathlib.Path(__file__).resolve().parents[1] / "optix_raytracer_b200" / "csrc" / "raycast.cu").read_text()
    m = re.search(r"vi = \*\(const BufView\*\)\(rec \+ (\d+)\), vp = \*\(const BufView\*\)\(rec \+ (\d+)\), vn = \*\(const BufView\*\)\(rec \+ (\d+)\)", src)
    assert m and tuple(int(x) for x in m.groups()) == (hl["off_indices"], hl["off_positions"], hl["off_normals"])


def test_whitted_struct_layouts_match_the_reference_headers():
    """MaterialData / Texture / Light / whitted::LaunchParams offsets measured on the reference headers against host.pack_material,
    host.WLaunchParams and the struct views csrc/whitted.cu static_asserts."""
    import struct as st
    torch = pytest.importorskip("torch")
    from optix_raytracer_b200 import host
    wl = KAT["whitted_layout"]
    assert (wl["MaterialData"], wl["Texture"], wl["Light"], wl["LaunchParams"]) == (240, 40, 36, 128)
    tex = {"index": 0, "texcoord": 1, "offset": [0.25, 0.5], "rotation": 0.0, "scale": [2.0, 3.0]}
    m = {"base_color": [0.1, 0.2, 0.3, 0.4], "metallic": 0.6, "roughness": 0.7, "base_color_tex": tex, "metallic_roughness_tex": None, "normal_tex": tex,
         "emissive_tex": None, "emissive_factor": [0.01, 0.02, 0.03], "alpha_mode": 1, "alpha_cutoff": 0.5, "double_sided": True}
    b = host.pack_material(m, {0: 0xABCDEF})
    assert len(b) == 240
    pbr = wl["pbr"]
    assert st.unpack_from("<4f", b, pbr + wl["pbr_base_color"]) == pytest.approx((0.1, 0.2, 0.3, 0.4))
    assert st.unpack_from("<f", b, pbr + wl["pbr_metallic"])[0] == pytest.approx(0.6) and st.unpack_from("<f", b, pbr + wl["pbr_roughness"])[0] == pytest.approx(0.7)
    assert st.unpack_from("<i", b, wl["alpha_mode"])[0] == 1 and st.unpack_from("<f", b, wl["alpha_cutoff"])[0] == 0.5
    assert st.unpack_from("<3f", b, wl["emissive_factor"]) == pytest.approx((0.01, 0.02, 0.03)) and b[wl["doubleSided"]] == 1
    for off in (wl["normal_tex"], pbr + wl["pbr_base_color_tex"]):
        assert st.unpack_from("<i", b, off + wl["tex_texcoord"])[0] == 1 and st.unpack_from("<Q", b, off + wl["tex_tex"])[0] == 0xABCDEF
        assert st.unpack_from("<2f", b, off + wl["tex_offset"]) == (0.25, 0.5) and st.unpack_from("<2f", b, off + wl["tex_scale"]) == (2.0, 3.0)
        assert st.unpack_from("<2f", b, off + wl["tex_rotation"]) == (0.0, 1.0)  # (sin, cos)
    assert st.unpack_from("<Q", b, wl["emissive_tex"] + wl["tex_tex"])[0] == 0 and st.unpack_from("<Q", b, pbr + wl["pbr_metallic_roughness_tex"] + wl["tex_tex"])[0] == 0
    P = host.WLaunchParams
    assert (P.subframe_index.offset, P.accum_buffer.offset, P.frame_buffer.offset, P.eye.offset, P.U.offset, P.lights_data.offset, P.miss_color.offset,
            P.handle.offset) == (wl["lp_subframe_index"], wl["lp_accum_buffer"], wl["lp_frame_buffer"], wl["lp_eye"], wl["lp_U"], wl["lp_lights"],
                                 wl["lp_miss_color"], wl["lp_handle"])
    assert (wl["light_type"], wl["light_point"] + wl["point_color"], wl["light_point"] + wl["point_intensity"], wl["light_point"] + wl["point_position"],
            wl["light_point"] + wl["point_falloff"]) == (0, 4, 16, 20, 32)


def test_optix_function_table_shim_answers_like_the_reference_expects():
    """optixQueryFunctionTable of optix_shim/libnvoptix.so.1 (csrc/optix_shim.cu): ABI 87 only, the 384-byte table of the reference's
    include/optix_function_table.h, every entry callable; the OptiX API structs it mirrors have the layout measured on the reference headers."""
    import ctypes as C
    from optix_raytracer_b200 import _lib
    shim = pathlib.Path(_lib.__file__).resolve().parent / "optix_shim" / "libnvoptix.so.1"
    assert shim.exists(), "csrc/Makefile builds it next to libb200rt.so"
    lib = C.CDLL(str(shim))
    q = lib.optixQueryFunctionTable
    q.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lay = KAT["optix_api_layout"]
    n = lay["sizeof_OptixFunctionTable"] // 8
    table = (C.c_void_p * n)()
    assert q(lay["OPTIX_ABI_VERSION"] + 1, 0, None, None, table, C.sizeof(table)) == 7801   # OPTIX_ERROR_UNSUPPORTED_ABI_VERSION
    assert q(lay["OPTIX_ABI_VERSION"], 0, None, None, table, C.sizeof(table) - 8) == 7802   # OPTIX_ERROR_FUNCTION_TABLE_SIZE_MISMATCH
    assert q(lay["OPTIX_ABI_VERSION"], 0, None, None, table, C.sizeof(table)) == 0
    assert all(table[i] for i in range(n))
    # optixGetErrorName / optixGetErrorString are the first two entries
    name = C.CFUNCTYPE(C.c_char_p, C.c_int)(table[0])
    assert name(7800) == b"OPTIX_ERROR_NOT_SUPPORTED" and name(0) == b"OPTIX_SUCCESS"
    # an entry without a restatement (optixDenoiserCreate, entry 40) says so
    assert C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)(table[40])(None, 0, None, None) == 7800
    out = (C.c_uint * 32)()
    k = lib.b200rt_optix_shim_layout(out, 32)
    assert list(out[:k]) == list(lay.values())


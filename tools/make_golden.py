#!/usr/bin/env python3
"""Generate tests/golden/kat.json from the REFERENCE's own code (oracle/_ref/libref_shim.so,
built by oracle/Makefile from headers under /root/reference, compiled in place).

Run in the build container only (needs /root/reference).  The JSON is committed; the GPU box and
the CPU test-suite read the JSON, never the reference.  Floats are stored as uint32 bit patterns so
the comparison is bit-exact.
"""
import ctypes, json, pathlib, random, struct, subprocess, sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "ref"])
ref = ctypes.CDLL(str(ROOT / "oracle" / "_ref" / "libref_shim.so"))
ref.ref_tea4.restype = ctypes.c_uint
ref.ref_tea4.argtypes = [ctypes.c_uint, ctypes.c_uint]
ref.ref_lcg.restype = ctypes.c_uint
ref.ref_lcg.argtypes = [ctypes.POINTER(ctypes.c_uint)]
ref.ref_rnd.restype = ctypes.c_float
ref.ref_rnd.argtypes = [ctypes.POINTER(ctypes.c_uint)]
ref.ref_wd_num_samples.restype = ctypes.c_int

def fbits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]

def main():
    rng = random.Random(20261018)
    out = {"generator": "tools/make_golden.py via oracle/_ref/libref_shim.so",
           "sources": ["SDK/cuda/random.h:30-67", "SDK/sutil/WorkDistribution.h:50-81",
                       "SDK/sutil/Camera.cpp:34-46", "SDK/optixPathTracer/optixPathTracer.cpp:439"]}
    # --- tea<4> + rnd streams
    pairs = [(0, 0), (1, 0), (0, 1), (589823, 0), (12345, 7), (589823, 15), (0xFFFFFFFF, 0xFFFFFFFF)]
    pairs += [(rng.randrange(1 << 32), rng.randrange(1 << 16)) for _ in range(57)]
    tea = []
    for v0, v1 in pairs:
        seed = ref.ref_tea4(v0, v1)
        st = ctypes.c_uint(seed)
        draws = [fbits(ref.ref_rnd(ctypes.byref(st))) for _ in range(8)]
        tea.append({"v0": v0, "v1": v1, "seed": seed, "rnd_bits": draws, "state_after": st.value})
    out["tea4_rnd"] = tea
    st = ctypes.c_uint(0)
    out["lcg_from_0"] = [[ref.ref_lcg(ctypes.byref(st)), st.value] for _ in range(16)]
    # --- StaticWorkDistribution
    wd = []
    for (w, h) in [(768, 768), (3840, 2160), (1920, 1080), (100, 37), (1, 1), (1040, 968)]:
        for n in (1, 2, 3, 4, 8):
            for gpu in sorted({0, n // 2, n - 1}):
                ns = ref.ref_wd_num_samples(w, h, n, gpu)
                samples = sorted({0, 1, 31, 32, 33, 1000 % ns, ns // 2, ns - 1} | {rng.randrange(ns) for _ in range(8)})
                pix = []
                for s in samples:
                    xy = (ctypes.c_int * 2)()
                    ref.ref_wd_sample_pixel(w, h, n, gpu, s, xy)
                    pix.append([s, xy[0], xy[1]])
                wd.append({"w": w, "h": h, "ngpu": n, "gpu": gpu, "num_samples": ns, "pixels": pix})
    out["work_distribution"] = wd
    # --- Camera::UVWFrame
    cams = [([278.0, 273.0, -900.0], [278.0, 273.0, 330.0], [0.0, 1.0, 0.0], 35.0, 1.0),
            ([278.0, 273.0, -900.0], [278.0, 273.0, 330.0], [0.0, 1.0, 0.0], 35.0, 3840.0 / 2160.0),
            ([0.0, 1.0, -10.0], [0.0, 0.1, 0.0], [0.0, 1.0, 0.00000073], 45.0, 1920.0 / 1080.0),
            ([1.5, 2.5, 3.5], [-0.25, 0.75, 0.1], [0.1, 0.9, 0.2], 60.0, 1.25)]
    cam_out = []
    for eye, lookat, up, fov, asp in cams:
        f3 = ctypes.c_float * 3
        uvw = (ctypes.c_float * 9)()
        ref.ref_camera_uvw(f3(*eye), f3(*lookat), f3(*up), ctypes.c_float(fov), ctypes.c_float(asp), uvw)
        cam_out.append({"eye": eye, "lookat": lookat, "up": up, "fov_y": fov, "aspect_bits": fbits(asp),
                        "uvw_bits": [fbits(x) for x in uvw]})
    out["camera_uvw"] = cam_out
    n = (ctypes.c_float * 3)()
    f3 = ctypes.c_float * 3
    ref.ref_light_normal(f3(0.0, 0.0, 105.0), f3(-130.0, 0.0, 0.0), n)
    out["cornell_light_normal_bits"] = [fbits(x) for x in n]
    # --- struct layouts of the launch ABI measured on the reference headers (oracle/ref_shim.cpp)
    o = (ctypes.c_int * 64)()
    k = ref.ref_hitgroup_layout(o)
    names = ["sizeof_HitGroupData", "sizeof_GeometryData", "sizeof_MaterialData", "off_indices", "off_positions", "off_normals", "off_texcoords0",
             "off_texcoords1", "off_colors", "sizeof_BufferView", "bv_off_data", "bv_off_count", "bv_off_byte_stride", "bv_off_elmt_byte_size",
             "off_material_data", "TRIANGLE_MESH"]
    assert k == len(names)
    out["hitgroup_layout"] = dict(zip(names, list(o[:k])))
    k = ref.ref_params_layout(o)
    names = ["pt_Params", "pt_Params_eye", "pt_Params_light", "pt_Params_handle", "pt_HitGroupData", "pt_HitGroupData_diffuse_color",
             "pt_HitGroupData_vertices", "pt_MissData", "mg_Params", "mg_Params_eye", "mg_Params_light", "mg_Params_handle",
             "mg_Params_sample_index_buffer", "mg_Params_device_idx", "rc_Params", "rc_Ray", "rc_Hit", "OptixBuildInput", "OptixInstance",
             "OptixShaderBindingTable", "OptixAccelBuildOptions", "OptixBuildInputTriangleArray"]
    assert k == len(names)
    out["params_layout"] = dict(zip(names, list(o[:k])))
    # --- imgui_test: layouts, and the host objects main.cpp:236-252 builds, as raw bytes written by the reference's own classes
    k = ref.ref_playground_layout(o)
    names = ["Params", "image_width", "image_height", "samples_per_frame", "camera", "dt", "dirty", "image", "film", "tfactor", "handle", "normals",
             "vertices", "mat_indices", "nmat_indices", "lights", "nlights", "materials", "nmaterials", "sizeof_Camera", "sizeof_LightVariant",
             "sizeof_DiffuseMaterial"]
    assert k == len(names)
    out["playground_layout"] = dict(zip(names, list(o[:k])))
    f3 = ctypes.c_float * 3
    cams = []
    for eye, up, lookat, aperture, fd, fov, ortho in [((0.0, 1.0, -10.0), (0.0, 0.0000073, 1.0), (0.0, 0.1, 0.0), 0.0, 1.0, 45.0, 0),
                                                       ((0.0, 1.0, -10.0), (0.0, 0.0000073, 1.0), (0.0, 0.1, 0.0), 0.05, 1.0, 45.0, 0),
                                                       ((0.3, 0.6, -1.2), (0.0, 1.0, 0.000073), (0.0, 0.1, 0.0), 0.02, 1.0, 50.0, 0),
                                                       ((0.3, 0.6, -1.2), (0.0, 1.0, 0.000073), (0.0, 0.1, 0.0), 0.0, 1.0, 50.0, 1)]:
        buf = (ctypes.c_uint8 * 92)()
        ref.ref_pg_camera(f3(*eye), f3(*up), f3(*lookat), ctypes.c_float(aperture), ctypes.c_float(fd), ctypes.c_float(fov), ortho, buf)
        rays = []
        for ix, iy, w, h, seed in [(0, 0, 64, 48, 1), (63, 47, 64, 48, 12345), (17, 30, 64, 48, 0xdeadbeef), (960, 540, 1920, 1080, 7)]:
            st = ctypes.c_uint(seed)
            og, dr = f3(), f3()
            ref.ref_pg_compute_ray(buf, ix, iy, w, h, ctypes.byref(st), og, dr)
            rays.append({"ix": ix, "iy": iy, "w": w, "h": h, "seed": seed, "seed_after": st.value, "org_bits": [fbits(x) for x in og], "dir_bits": [fbits(x) for x in dr]})
        cams.append({"eye": eye, "up": up, "lookat": lookat, "aperture": aperture, "fd": fd, "fov": fov, "ortho": ortho, "bytes": bytes(buf).hex(), "rays": rays})
    out["playground_cameras"] = cams
    lights = []
    for kind, a, lumi, scalar in [(2, (0.0, 2.0, 0.0), (0.1, 0.08, 0.08), 0.1), (2, (2.0, 2.0, 0.0), (0.1, 0.08, 0.08), 0.1), (2, (2.0, 2.0, 2.0), (0.1, 0.08, 0.08), 0.1),
                                  (1, (-1.0, 1.0, -1.0), (0.1, 0.1, 0.1), 0.05), (0, (0.5, 1.5, -0.5), (0.2, 0.3, 0.4), 0.0)]:
        buf = (ctypes.c_uint8 * 44)()
        ref.ref_pg_light(kind, f3(*a), f3(*lumi), ctypes.c_float(scalar), buf)
        st = ctypes.c_uint(4242)
        wi, lm = f3(), f3()
        ref.ref_pg_light_eval(buf, f3(0.1, 0.05, -0.2), ctypes.byref(st), wi, lm)
        lights.append({"kind": kind, "a": a, "lumi": lumi, "scalar": scalar, "bytes": bytes(buf).hex(), "p": [0.1, 0.05, -0.2], "seed": 4242,
                       "seed_after": st.value, "wi_bits": [fbits(x) for x in wi], "lumi_bits": [fbits(x) for x in lm]})
    out["playground_lights"] = lights
    k = ref.ref_whitted_layout(o)
    names = ["MaterialData", "type", "normal_tex", "alpha_mode", "alpha_cutoff", "emissive_factor", "emissive_tex", "doubleSided", "pbr", "pbr_base_color",
             "pbr_metallic", "pbr_roughness", "pbr_base_color_tex", "pbr_metallic_roughness_tex", "Texture", "tex_texcoord", "tex_tex", "tex_offset",
             "tex_rotation", "tex_scale", "Light", "light_type", "light_point", "point_color", "point_intensity", "point_position", "point_falloff",
             "LaunchParams", "lp_subframe_index", "lp_accum_buffer", "lp_frame_buffer", "lp_eye", "lp_U", "lp_lights", "lp_miss_color", "lp_handle"]
    assert k == len(names)
    out["whitted_layout"] = dict(zip(names, list(o[:k])))
    k = ref.ref_optix_api_layout(o)
    names = ["OPTIX_ABI_VERSION", "sizeof_OptixFunctionTable", "sizeof_OptixDeviceContextOptions", "sizeof_OptixProgramGroupDesc", "off_single_entryFunctionName",
             "off_hitgroup_entryFunctionNameCH", "off_hitgroup_entryFunctionNameAH", "off_hitgroup_entryFunctionNameIS", "sizeof_OptixPipelineCompileOptions",
             "sizeof_OptixStackSizes", "KIND_RAYGEN", "KIND_MISS", "KIND_HITGROUP", "PROPERTY_RTCORE_VERSION"]
    assert k == len(names)
    out["optix_api_layout"] = dict(zip(names, list(o[:k])))
    p = ROOT / "tests" / "golden" / "kat.json"
    p.write_text(json.dumps(out) + "\n")
    print("wrote", p, p.stat().st_size, "bytes")

if __name__ == "__main__":
    sys.exit(main())

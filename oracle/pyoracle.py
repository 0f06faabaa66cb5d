"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE — the checker, never the product).

Import is allowed only from tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference legs).  The library is built by `make -C oracle` (done by __graft_entry__.build()).
"""
import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


def build(force=False):
    so = _HERE / "liboracle.so"
    src = _HERE / "oracle.cpp"
    if force or not so.exists() or (src.exists() and src.stat().st_mtime > so.stat().st_mtime):
        subprocess.check_call(["make", "-s", "-C", str(_HERE), str(so)])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        u32, i32, i64, f32, vp = C.c_uint32, C.c_int32, C.c_int64, C.c_float, C.c_void_p
        L.orc_tea4.restype = u32; L.orc_tea4.argtypes = [u32, u32]
        L.orc_lcg.restype = u32; L.orc_lcg.argtypes = [C.POINTER(u32)]
        L.orc_rnd.restype = f32; L.orc_rnd.argtypes = [C.POINTER(u32)]
        L.orc_sincos.argtypes = [f32, C.POINTER(f32), C.POINTER(f32)]
        L.orc_wd_num_samples.restype = i32; L.orc_wd_num_samples.argtypes = [i32, i32, i32]
        L.orc_wd_sample_pixel.argtypes = [i32, i32, i32, i32, i32, C.POINTER(i32)]
        L.orc_camera_uvw.argtypes = [vp, vp, vp, f32, f32, vp]
        L.orc_make_color.argtypes = [vp, i32, vp]
        L.orc_scene_create.restype = vp; L.orc_scene_create.argtypes = [vp, i64, vp]
        L.orc_scene_add_instance.argtypes = [vp, vp]
        L.orc_scene_add_instance_ex.argtypes = [vp, vp, u32, u32]
        L.orc_cull_word.argtypes = [u32, u32]
        L.orc_cull_word.restype = u32
        L.orc_scene_set_triangle_flags.argtypes = [vp, vp]
        L.orc_scene_set_brute.argtypes = [vp, i32]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_invert34.argtypes = [vp, vp]
        L.orc_trace.argtypes = [vp, vp, i64, vp, i32, u32, i32, vp]
        L.orc_raycast_ortho_scalars.argtypes = [vp, vp, i32, i32, f32, vp]
        L.orc_raycast_create_rays.argtypes = [vp, i32, i32, f32, f32, f32, f32, f32]
        L.orc_raycast_translate.argtypes = [vp, i64, vp]
        L.orc_raycast_hits.argtypes = [vp, vp, i64, vp, vp, vp, i32]
        L.orc_raycast_shade.argtypes = [vp, i64, vp]
        L.orc_synth_mesh.argtypes = [C.c_uint64, u32, vp, vp, i32]
        L.orc_scene_set_geometry_flags.argtypes = [vp, u32]
        L.orc_playground.restype = C.c_uint64
        L.orc_playground.argtypes = [vp, vp, vp, i32, vp, vp, vp, i32, i32, u32, u32, i32, vp, vp, i32, i32, i32]
        L.orc_playground_scene.restype = C.c_uint64
        L.orc_playground_scene.argtypes = [u32, u32, vp, vp, vp, i32]
        L.orc_whitted.restype = C.c_uint64
        L.orc_whitted.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32]
        L.orc_pathtrace.restype = C.c_uint64
        L.orc_pathtrace.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ncores():
    return len(os.sched_getaffinity(0))


class WhittedParams(C.Structure):  # oracle.cpp: orc_whitted_params
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("subframe_index", C.c_uint32), ("eye", C.c_float * 3), ("U", C.c_float * 3),
                ("V", C.c_float * 3), ("W", C.c_float * 3), ("miss_color", C.c_float * 3), ("base_color", C.c_float * 4), ("metallic", C.c_float),
                ("roughness", C.c_float), ("emissive", C.c_float * 3), ("nlights", C.c_int32), ("alpha_mode", C.c_int32)]


class PTParams(C.Structure):
    _fields_ = [("subframe_index", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32),
                ("samples_per_launch", C.c_int32),
                ("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3),
                ("light_corner", C.c_float * 3), ("light_v1", C.c_float * 3), ("light_v2", C.c_float * 3),
                ("light_normal", C.c_float * 3), ("light_emission", C.c_float * 3), ("bg", C.c_float * 3),
                ("nmat", C.c_int32), ("mode", C.c_int32), ("groups", C.c_int32)]


def tea4(v0, v1):
    return lib().orc_tea4(v0, v1)


def rnd_stream(seed, n):
    st = C.c_uint32(seed)
    vals = [lib().orc_rnd(C.byref(st)) for _ in range(n)]
    return np.array(vals, dtype=np.float32), st.value


def camera_uvw(eye, lookat, up, fovy, aspect):
    out = np.zeros(9, np.float32)
    lib().orc_camera_uvw(_p(_f32(eye)), _p(_f32(lookat)), _p(_f32(up)), np.float32(fovy), np.float32(aspect), _p(out))
    return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()


def wd_num_samples(w, h, ngpu):
    return lib().orc_wd_num_samples(w, h, ngpu)


def wd_sample_pixel(w, h, ngpu, gpu, sample):
    xy = (C.c_int32 * 2)()
    lib().orc_wd_sample_pixel(w, h, ngpu, gpu, sample, xy)
    return xy[0], xy[1]


def make_color(rgb):
    rgb = _f32(rgb).reshape(-1, 3)
    out = np.zeros((rgb.shape[0], 4), np.uint8)
    lib().orc_make_color(_p(rgb), rgb.shape[0], _p(out))
    return out


class Scene:
    """Triangle soup (ntri,3,3) float32 in object space + optional per-triangle SBT offsets and instances.
    instances: 3x4 transforms, or (transform, OptixInstanceFlags, visibilityMask) tuples.  geom_flags: OptixGeometryFlags of the build
    input — one value, or one per triangle (the flags of the SBT record the triangle belongs to)."""

    def __init__(self, tris, sbt=None, instances=(), geom_flags=None):
        self.tris = _f32(tris).reshape(-1, 9)
        self.sbt = None if sbt is None else np.ascontiguousarray(sbt, dtype=np.uint32)
        self.h = lib().orc_scene_create(_p(self.tris), self.tris.shape[0], _p(self.sbt))
        for m in instances:
            if isinstance(m, tuple):
                lib().orc_scene_add_instance_ex(self.h, _p(_f32(m[0]).reshape(12)), int(m[1]), int(m[2]) if len(m) > 2 else 1)
            else:
                lib().orc_scene_add_instance(self.h, _p(_f32(m).reshape(12)))
        if geom_flags is not None:  # OptixGeometryFlags of the build input (4 = DISABLE_TRIANGLE_FACE_CULLING)
            if np.ndim(geom_flags) == 0:
                lib().orc_scene_set_geometry_flags(self.h, int(geom_flags))
            else:
                gf = np.ascontiguousarray(geom_flags, dtype=np.uint8)
                assert gf.shape[0] == self.tris.shape[0]
                lib().orc_scene_set_triangle_flags(self.h, _p(gf))

    def set_brute(self, brute):
        lib().orc_scene_set_brute(self.h, int(bool(brute)))

    def __del__(self):
        try:
            lib().orc_scene_destroy(self.h)
        except Exception:
            pass

    def trace(self, rays, any_hit=False, ray_flags=0, threads=None, stats=False):
        """rays (n,8) float32 {o,tmin,d,tmax}.  Returns dict(t,prim,inst,b1,b2) or occluded bool array."""
        rays = _f32(rays).reshape(-1, 8)
        n = rays.shape[0]
        out = np.zeros((n, 5), np.uint32)
        st = np.zeros(2, np.uint64) if stats else None
        lib().orc_trace(self.h, _p(rays), n, _p(out), int(any_hit), ray_flags, threads or ncores(), _p(st))
        if any_hit:
            res = {"occluded": out[:, 0].astype(bool)}
        else:
            res = {"t": out[:, 0].view(np.float32).copy(), "prim": out[:, 1].copy(), "inst": out[:, 2].copy(),
                   "b1": out[:, 3].view(np.float32).copy(), "b2": out[:, 4].view(np.float32).copy()}
        if stats:
            res["node_visits"], res["tri_tests"] = int(st[0]), int(st[1])
        return res

    def raycast_hits(self, rays, normals=None, threads=None):
        rays = _f32(rays).reshape(-1, 8)
        n = rays.shape[0]
        nn = None if normals is None else _f32(normals).reshape(-1, 9)
        hits = np.zeros((n, 4), np.float32)
        ext = np.zeros((n, 5), np.uint32)
        lib().orc_raycast_hits(self.h, _p(rays), n, _p(nn), _p(hits), _p(ext), threads or ncores())
        return hits, ext

    def set_geometry_flags(self, gflags):
        """OptixGeometryFlags of the build input (default 1 = DISABLE_ANYHIT; imgui_test builds with 0)."""
        lib().orc_scene_set_geometry_flags(self.h, int(gflags))

    def playground(self, camera92, lights44, materials, normals, mat_indices, width, height, spf, dt, dirty, film=None, rows=None, threads=None,
                   want_image=True):
        """imgui_test frame (oracle.cpp: orc_playground).  camera92 / lights44: raw bytes of the reference's Camera / LightVariant[] objects.
        Returns (film (h,w,3) f32, image (h,w,4) u8 or None, rays traced)."""
        cam = np.frombuffer(bytes(camera92), np.uint8).copy()
        lights = np.frombuffer(bytes(lights44), np.uint8).copy()
        nl = lights.size // 44
        mats = _f32(materials).reshape(-1)
        nrm = _f32(normals).reshape(-1)
        mi = np.ascontiguousarray(mat_indices, dtype=np.int32)
        if film is None:
            film = np.zeros((height, width, 3), np.float32)
        image = np.zeros((height, width, 4), np.uint8) if want_image else None
        y0, y1 = rows or (0, height)
        n = lib().orc_playground(self.h, _p(cam), _p(lights), nl, _p(mats), _p(nrm), _p(mi), width, height, spf, dt, int(bool(dirty)), _p(film), _p(image),
                                 y0, y1, threads or ncores())
        return film, image, int(n)

    def whitted(self, params, lights36, normals=None, accum=None, rows=None, threads=None):
        """optixMeshViewer frame for an untextured material (params.alpha_mode 2 = BLEND) (oracle.cpp: orc_whitted).  Returns (accum (h,w,4), frame (h,w,4) u8, rays)."""
        w, h = params.width, params.height
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        frame = np.zeros((h, w, 4), np.uint8)
        lights = np.frombuffer(bytes(lights36), np.uint8).copy()
        params.nlights = lights.size // 36
        nn = None if normals is None else _f32(normals).reshape(-1)
        y0, y1 = rows or (0, h)
        n = lib().orc_whitted(self.h, C.byref(params), _p(nn), _p(lights), _p(accum), _p(frame), y0, y1, threads or ncores())
        return accum, frame, int(n)

    def pathtrace(self, params, emission, diffuse, accum=None, region=None, threads=None, want_frame=True):
        """params: PTParams.  Returns (accum (h,w,4) f32, frame (h,w,4) u8, segments)."""
        w, h = params.width, params.height
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        frame = np.zeros((h, w, 4), np.uint8) if want_frame else None
        x0, y0, x1, y1 = region or (0, 0, w, h)
        em, df = _f32(emission).reshape(-1), _f32(diffuse).reshape(-1)
        segs = lib().orc_pathtrace(self.h, C.byref(params), _p(em), _p(df), _p(accum), _p(frame), x0, y0, x1, y1,
                                   threads or ncores())
        return accum, frame, int(segs)


def raycast_ortho_scalars(bbmin, bbmax, width, height, padding=0.05):
    out = np.zeros(5, np.float32)
    lib().orc_raycast_ortho_scalars(_p(_f32(bbmin)), _p(_f32(bbmax)), width, height, np.float32(padding), _p(out))
    return out  # x0, y0, z, dx, dy


def raycast_create_rays(width, height, x0, y0, z, dx, dy):
    rays = np.zeros((height * width, 8), np.float32)
    lib().orc_raycast_create_rays(_p(rays), width, height, np.float32(x0), np.float32(y0), np.float32(z),
                                  np.float32(dx), np.float32(dy))
    return rays


def raycast_translate(rays, off):
    rays = _f32(rays).copy()
    lib().orc_raycast_translate(_p(rays), rays.shape[0], _p(_f32(off)))
    return rays


def raycast_shade(hits):
    hits = _f32(hits).reshape(-1, 4)
    img = np.zeros((hits.shape[0], 3), np.float32)
    lib().orc_raycast_shade(_p(hits), hits.shape[0], _p(img))
    return img


def invert34(m):
    out = np.zeros(12, np.float32)
    lib().orc_invert34(_p(_f32(m).reshape(12)), _p(out))
    return out


def synth_mesh(total, seed=0, threads=None):
    """(total,3,3) float32 triangles and (total,) uint32 material indices of the synthetic scene."""
    verts = np.zeros((total, 9), np.float32)
    mats = np.zeros(total, np.uint32)
    lib().orc_synth_mesh(total, seed, _p(verts), _p(mats), threads or ncores())
    return verts.reshape(total, 3, 3), mats


def playground_scene(rows, seed=0, threads=None):
    """Stand-in scene of imgui_test (playground.cu: pg_scene_kernel): (T,3,3) vertices, (T,3,3) normals, (T,) int32 materials."""
    total = int(lib().orc_playground_scene(rows, seed, None, None, None, 1))
    verts = np.zeros((total, 9), np.float32)
    nrm = np.zeros((total, 9), np.float32)
    mats = np.zeros(total, np.int32)
    lib().orc_playground_scene(rows, seed, _p(verts), _p(nrm), _p(mats), threads or ncores())
    return verts.reshape(total, 3, 3), nrm.reshape(total, 3, 3), mats

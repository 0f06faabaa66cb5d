// TEST INFRASTRUCTURE — not product code.  Second translation unit of oracle/_ref/libref_shim.so: the reference's imgui_test
// host classes compiled in place (their `Params` would clash with the other samples' in ref_shim.cpp).
#include <cuda_runtime.h>
#include <cstddef>
#include <cstring>
#include <optix_types.h>

// ---- imgui_test ("playground") host objects, built through the reference's own classes -------------------------------------
// (SDK/imgui_test/camera.h, light.h, volumetric_light.h, directional_light.h, point_light.h, diffuse.h, optixTriangle.h)
#include <imgui_test/optixTriangle.h>

extern "C" {

int ref_playground_layout(int* out)
{
    int n = 0;
    out[n++] = (int)sizeof(Params);
    out[n++] = (int)offsetof(Params, image_width); out[n++] = (int)offsetof(Params, image_height);
    out[n++] = (int)offsetof(Params, samples_per_frame); out[n++] = (int)offsetof(Params, camera); out[n++] = (int)offsetof(Params, dt);
    out[n++] = (int)offsetof(Params, dirty); out[n++] = (int)offsetof(Params, image); out[n++] = (int)offsetof(Params, film);
    out[n++] = (int)offsetof(Params, tfactor); out[n++] = (int)offsetof(Params, handle); out[n++] = (int)offsetof(Params, normals);
    out[n++] = (int)offsetof(Params, vertices); out[n++] = (int)offsetof(Params, mat_indices); out[n++] = (int)offsetof(Params, nmat_indices);
    out[n++] = (int)offsetof(Params, lights); out[n++] = (int)offsetof(Params, nlights); out[n++] = (int)offsetof(Params, materials);
    out[n++] = (int)offsetof(Params, nmaterials);
    out[n++] = (int)sizeof(Camera); out[n++] = (int)sizeof(LightVariant); out[n++] = (int)sizeof(DiffuseMaterial);
    return n;
}

// Camera set up exactly as main.cpp:236-243 does (setters, then compute_uvw); raw object bytes out
void ref_pg_camera(const float* eye, const float* up, const float* lookat, float aperture, float fd, float fov, int ortho, void* out92)
{
    Camera cam{};
    cam.set_eye(make_float3(eye[0], eye[1], eye[2]));
    cam.set_up(make_float3(up[0], up[1], up[2]));
    cam.set_lookat(make_float3(lookat[0], lookat[1], lookat[2]));
    cam.set_aperture(aperture);
    cam.set_fd(fd);
    cam.set_fov(fov);
    cam.set_ortho(ortho != 0);
    cam.compute_uvw();
    memcpy(out92, (const void*)&cam, sizeof cam);
}

// kind 0 = PointLight(position, lumi), 1 = DirectionalLight(direction, lumi, jitter), 2 = VolumetricLight(position, radius, lumi)
void ref_pg_light(int kind, const float* a, const float* lumi, float scalar, void* out44)
{
    const float3 A = make_float3(a[0], a[1], a[2]), Lm = make_float3(lumi[0], lumi[1], lumi[2]);
    if (kind == 0) { LightVariant l(PointLight(A, Lm)); memcpy(out44, (const void*)&l, sizeof l); }
    else if (kind == 1) { LightVariant l(DirectionalLight(A, Lm, scalar)); memcpy(out44, (const void*)&l, sizeof l); }
    else { LightVariant l(VolumetricLight(A, scalar, Lm)); memcpy(out44, (const void*)&l, sizeof l); }
}

// Camera::compute_ray and LightVariant::wi / lumi as the host compiler evaluates them (no contraction: agree with the
// contract's fma placement to within an ulp or two, used as a sanity bound, not bit-exactly)
void ref_pg_compute_ray(const void* cam92, unsigned ix, unsigned iy, unsigned w, unsigned h, unsigned* seed, float* org, float* dir)
{
    Camera cam{};
    memcpy((void*)&cam, cam92, sizeof cam);
    float3 o, d;
    cam.compute_ray(make_uint3(ix, iy, 0), make_uint3(w, h, 1), o, d, *seed);
    org[0] = o.x; org[1] = o.y; org[2] = o.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}
void ref_pg_light_eval(const void* light44, const float* p, unsigned* seed, float* wi, float* lumi)
{
    LightVariant l(PointLight{});
    memcpy((void*)&l, light44, sizeof l);
    const float3 w = l.wi(make_float3(p[0], p[1], p[2]), *seed), m = l.lumi();
    wi[0] = w.x; wi[1] = w.y; wi[2] = w.z; lumi[0] = m.x; lumi[1] = m.y; lumi[2] = m.z;
}

}  // extern "C"

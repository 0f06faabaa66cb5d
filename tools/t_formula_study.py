#!/usr/bin/env python3
"""Offline study of the hit distance: which fp32 formula for t reproduces what OptiX reports?  Reads gpurun_out/t_compare.npz
(tools/dump_t_compare.py: same rays through OptiX's built-in triangles and through b200rt) and evaluates candidate formulas in numpy
fp32 (fma = exact product and sum in fp64, rounded once to fp32) on the (ray, triangle) pairs both agree on.
    python tools/t_formula_study.py [gpurun_out/t_compare.npz]"""
import sys
import numpy as np

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def dot(a, b, fused):
    if fused:  # nvcc contracts a.x*b.x + a.y*b.y + a.z*b.z into fma(a.z, b.z, fma(a.y, b.y, a.x*b.x))
        return fma(a[:, 2], b[:, 2], fma(a[:, 1], b[:, 1], (a[:, 0] * b[:, 0]).astype(f32)))
    return ((a[:, 0] * b[:, 0]).astype(f32) + (a[:, 1] * b[:, 1]).astype(f32) + (a[:, 2] * b[:, 2]).astype(f32)).astype(f32)


def cross(a, b, fused):
    def comp(i, j):
        if fused:
            return fma(a[:, i], b[:, j], -(a[:, j] * b[:, i]).astype(f32))
        return ((a[:, i] * b[:, j]).astype(f32) - (a[:, j] * b[:, i]).astype(f32)).astype(f32)
    return np.stack([comp(1, 2), comp(2, 0), comp(0, 1)], axis=1).astype(f32)


def ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def report(name, t, ref):
    u = ulps(t.astype(f32), ref)
    print(f"  {name:58s} identical {(u == 0).mean():.4f}  <=1 {(u <= 1).mean():.4f}  <=2 {(u <= 2).mean():.4f}  <=4 {(u <= 4).mean():.4f}  max {u.max()}")
    return u


def study(label, o, d, v0, v1, v2, t_optix, t_ours):
    print(f"{label}: {o.shape[0]} (ray, triangle) pairs")
    report("b200rt (watertight, sheared space, T / det)", t_ours, t_optix)
    for fused in (False, True):
        tag = "fma" if fused else "mul+add"
        e1, e2 = (v1 - v0).astype(f32), (v2 - v0).astype(f32)
        # Moeller-Trumbore
        p = cross(d, e2, fused); det = dot(e1, p, fused); s = (o - v0).astype(f32); q = cross(s, e1, fused)
        t = (dot(e2, q, fused) / det).astype(f32)
        report(f"Moeller-Trumbore, {tag}, t = dot(e2, q) / det", t, t_optix)
        t = (dot(e2, q, fused) * (f32(1.0) / det).astype(f32)).astype(f32)
        report(f"Moeller-Trumbore, {tag}, t = dot(e2, q) * (1 / det)", t, t_optix)
        # plane equation
        n = cross(e1, e2, fused)
        t = (dot(n, (v0 - o).astype(f32), fused) / dot(n, d, fused)).astype(f32)
        report(f"plane, {tag}, t = dot(N, v0 - o) / dot(N, d)", t, t_optix)
        n2 = cross(e2, e1, fused)
        t = (dot(n2, (o - v0).astype(f32), fused) / -dot(n2, d, fused)).astype(f32) if False else t
        # plane through each vertex
        for k, vk in enumerate((v1, v2)):
            t = (dot(n, (vk - o).astype(f32), fused) / dot(n, d, fused)).astype(f32)
            report(f"plane, {tag}, t = dot(N, v{k + 1} - o) / dot(N, d)", t, t_optix)
    # error of ours vs t magnitude
    u = ulps(t_ours, t_optix)
    for lo, hi in ((0, 1), (1, 10), (10, 100), (100, 1000), (1000, 1e30)):
        m = (t_optix >= lo) & (t_optix < hi)
        if m.any():
            print(f"  t in [{lo:g}, {hi:g}): {int(m.sum()):8d} hits, b200rt <=1 ulp {(u[m] <= 1).mean():.4f}, <=4 ulp {(u[m] <= 4).mean():.4f}, max {u[m].max()}")


def main():
    z = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/t_compare.npz")
    rays = z["cornell_rays"]; verts = z["cornell_verts"].reshape(-1, 3, 3)
    tb, to, pb, po = z["cornell_b_t"], z["cornell_o_t"], z["cornell_b_prim"], z["cornell_o_prim"]
    hit = (tb >= 0) & (to >= 0) & (pb == po)
    tri = verts[po[hit].astype(np.int64)]
    study("Cornell (GAS)", rays[hit, 0:3].astype(f32), rays[hit, 4:7].astype(f32), tri[:, 0], tri[:, 1], tri[:, 2], to[hit], tb[hit])
    rays = z["duck_rays"]; tris = z["duck_tris"]; M = z["duck_xform"]
    tb, to, pb, po = z["duck_b_t"], z["duck_o_t"], z["duck_b_prim"], z["duck_o_prim"]
    hit = (tb >= 0) & (to >= 0) & (pb == po)
    tri = tris[po[hit].astype(np.int64)]
    # object-space ray: inverse of the instance transform (uniform scale + rotation + translation in the Duck)
    Minv = np.linalg.inv(M.astype(np.float64)).astype(f32)
    o = (rays[hit, 0:3] @ Minv[:3, :3].T + Minv[:3, 3]).astype(f32)
    d = (rays[hit, 4:7] @ Minv[:3, :3].T).astype(f32)
    study("Duck (IAS; object-space ray by a numpy inverse: indicative only)", o, d, tri[:, 0], tri[:, 1], tri[:, 2], to[hit], tb[hit])


if __name__ == "__main__":
    main()

"""Shared helpers for the parity tests (scene fixtures, oracle scene construction)."""
import ctypes as C
import pathlib

import numpy as np

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


def duck_scene(textured=True):
    """The reference's Duck.gltf as a host.load_gltf-shaped dict (from tests/golden/duck_mesh.npz): geometry, instance, camera, the
    material (metallic 0, roughness 1, base-colour texture) and a 128x128 stand-in of its texture."""
    z = np.load(GOLDEN / "duck_mesh.npz")
    prim = {"positions": z["positions"], "normals": z["normals"], "indices": z["indices"], "material": 0, "texcoords": [z["texcoords0"], None]}
    mesh = {"primitives": [prim], "aabb": (z["aabb_lo"], z["aabb_hi"])}
    inst = {"transform": z["transform"], "mesh": 0, "world_aabb": (z["world_lo"], z["world_hi"])}
    tex = {"index": 0, "texcoord": 0, "offset": [0.0, 0.0], "rotation": 0.0, "scale": [1.0, 1.0]}
    mat = {"base_color": [1.0, 1.0, 1.0, 1.0], "metallic": 0.0, "roughness": 1.0, "base_color_tex": tex if textured else None,
           "metallic_roughness_tex": None, "normal_tex": None, "emissive_tex": None, "emissive_factor": [0.0, 0.0, 0.0], "alpha_mode": 0,
           "alpha_cutoff": 0.5, "double_sided": False}
    return {"meshes": [mesh], "instances": [inst], "cameras": [{"eye": z["cam_eye"], "up": z["cam_up"], "fov_y": float(z["cam_fov_y"])}],
            "materials": [mat], "images": [np.ascontiguousarray(z["texture_rgba8"])], "textures": [{"sampler": 0, "source": 0}],
            "samplers": [{"magFilter": 9729, "minFilter": 9986, "wrapS": 10497, "wrapT": 10497}]}


def duck_alpha_scene(alpha_mode, cutoff=0.5):
    """duck_scene() with the base-colour texture's alpha channel turned into a pattern of holes (0), solid (255) and smooth ramps in between
    — the role DuckHole.gltf's DuckHole.png plays in the reference (alphaMode MASK, alphaCutoff 0.5) — and the material's alpha mode set
    (1 = MASK, 2 = BLEND)."""
    sc = duck_scene()
    img = sc["images"][0].copy()
    h, w = img.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = (128 + 127 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.int64)
    holes = ((xx // 16 + yy // 16) % 3 == 0)
    solid = ((xx // 16 + yy // 16) % 3 == 1)
    img[..., 3] = np.where(holes, 0, np.where(solid, 255, ramp)).astype(np.uint8)
    sc["images"][0] = np.ascontiguousarray(img)
    sc["materials"][0].update({"alpha_mode": alpha_mode, "alpha_cutoff": cutoff})
    return sc


def deindex(prim):
    """(ntri,3,3) object-space triangles and per-triangle vertex normals of one primitive group."""
    idx = prim["indices"].astype(np.int64).reshape(-1, 3)
    tris = prim["positions"][idx]
    nrm = None if prim["normals"] is None else prim["normals"][idx]
    return tris.astype(np.float32), None if nrm is None else nrm.astype(np.float32)


def oracle_pt_params(orc, params, mode, nmat=4):
    """orc.PTParams from a host.PTParams / host.MGParams."""
    p = orc.PTParams()
    p.subframe_index, p.width, p.height = params.subframe_index, params.width, params.height
    p.samples_per_launch, p.nmat, p.mode = params.samples_per_launch, nmat, mode
    for k in ("eye", "U", "V", "W"):
        setattr(p, k, getattr(params, k))
    for k in ("corner", "v1", "v2", "normal", "emission"):
        setattr(p, "light_" + k, getattr(params.light, k))
    p.bg = (C.c_float * 3)(0.0, 0.0, 0.0)
    return p


def random_rays(rng, n, lo, hi, tmin=0.0, tmax=1e16):
    """Rays with origins in a padded box and directions towards random points of the box (plus axis-aligned
    and degenerate-component directions, which exercise the clamped-direction box tests)."""
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    ext = hi - lo
    o = (lo - 0.3 * ext + rng.random((n, 3), dtype=np.float32) * 1.6 * ext).astype(np.float32)
    tgt = (lo + rng.random((n, 3), dtype=np.float32) * ext).astype(np.float32)
    d = tgt - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    k = n // 8
    axes = np.eye(3, dtype=np.float32)
    d[:k] = axes[rng.integers(0, 3, k)] * rng.choice(np.array([-1.0, 1.0], np.float32), (k, 1))
    zero_one = rng.integers(0, 3, k)
    d[k:2 * k][np.arange(k), zero_one] = 0.0
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = tmin
    rays[:, 4:7] = d.astype(np.float32)
    rays[:, 7] = tmax
    return rays


def edge_case_scene(seed=23):
    """Triangles and rays of test_edge_case_rays_and_geometry_vs_oracle (tests/test_gpu_parity.py): a coplanar grid with shared edges,
    a fan of slivers, a small grid a million units from the origin; rays at vertices, edge midpoints, along diagonals, in the plane of
    the grid, with degenerate intervals, plus random ones."""
    rng = np.random.default_rng(seed)
    G = 24
    xs = np.arange(G + 1, dtype=np.float32) * np.float32(0.25)
    tris = []
    for i in range(G):          # coplanar grid at z = 1, two triangles per cell, shared edges everywhere
        for j in range(G):
            a, b, c, d = (xs[i], xs[j], 1), (xs[i + 1], xs[j], 1), (xs[i + 1], xs[j + 1], 1), (xs[i], xs[j + 1], 1)
            tris += [[a, b, c], [a, c, d]]
    fan_c = (3.0, 3.0, 2.0)      # fan of slivers around one vertex
    for k in range(64):
        a0, a1 = 2 * np.pi * k / 64, 2 * np.pi * (k + 1) / 64
        tris.append([fan_c, (3 + 2 * np.cos(a0), 3 + 2 * np.sin(a0), 2.0 + 1e-3 * k), (3 + 2 * np.cos(a1), 3 + 2 * np.sin(a1), 2.0 + 1e-3 * (k + 1))])
    far = np.float32(1.0e6)      # the same little grid a million units away (coarse float spacing: 0.0625)
    for i in range(4):
        for j in range(4):
            a, b, c, d = (far + i, far + j, far), (far + i + 1, far + j, far), (far + i + 1, far + j + 1, far), (far + i, far + j + 1, far)
            tris += [[a, b, c], [a, c, d]]
    tris = np.asarray(tris, np.float32)
    rays = []
    for i in range(G + 1):       # straight down at every grid vertex, and at edge midpoints
        for j in range(G + 1):
            rays.append([xs[i], xs[j], 5, 0, 0, 0, -1, 100])
            if i < G:
                rays.append([(xs[i] + xs[i + 1]) / 2, xs[j], 5, 0, 0, 0, -1, 100])
            if i < G and j < G:  # along the cell diagonal (the edge shared by the two triangles of a cell), obliquely
                rays.append([xs[i] + 0.125, xs[j] + 0.125, 5, 0, 0.25, 0.25, -1, 100])
    for k in range(2000):        # rays lying in the plane z = 1 (parallel to the grid), and axis-parallel ones through the fan
        y = np.float32(rng.random() * 6)
        rays.append([-1, y, 1, 0, 1, 0, 0, 100])
        rays.append([3, 3, -1, 0, 0, 0, 1, 100])
        rays.append([np.float32(rng.random() * 6), y, 3, 0, 0, 0, -1, np.inf])
    rays += [[1, 1, 5, 0, 0, 0, -1, 4.0], [1, 1, 5, 4.0, 0, 0, -1, 4.0], [1, 1, 5, 4.5, 0, 0, -1, 4.0], [1, 1, 5, 0, 0, 0, -1, -1.0],
             [1, 1, 1, 0, 0, 0, -1, 100], [1, 1, 1, 0, 0, 0, 1, 100], [0.3, 0.3, 1, 0, 1, 1, 0, 100]]
    for i in range(5):           # at the far grid: vertices, edges, interior
        for j in range(5):
            rays.append([far + i, far + j, far + 10, 0, 0, 0, -1, 100])
            rays.append([far + i + 0.5, far + j + 0.25, far + 10, 0, 0, 0, -1, 100])
    rnd = random_rays(rng, 20000, [-1, -1, 0], [7, 7, 4])
    rays = np.concatenate([np.asarray(rays, np.float32), rnd]).astype(np.float32)
    return tris, rays


def decode_gas(blob):
    """Decode a GAS blob (optix_raytracer_b200/csrc/accel.h) downloaded from the device into numpy views."""
    hdr = np.frombuffer(blob[:128].tobytes(), dtype=np.uint32)
    assert hdr[0] == 0x54523242 and hdr[1] == 1, "not a GAS blob"
    num_tris, num_nodes = int(hdr[2]), int(hdr[3])
    nodes_off, tris_off, total = [int(x) for x in np.frombuffer(blob[16:40].tobytes(), dtype=np.uint64)]
    bounds = np.frombuffer(blob[40:64].tobytes(), dtype=np.float32)
    node_bytes = int(hdr[22]) or 80   # AccelHeader::node_bytes: 80 = Node8 (8-bit boxes), 224 = Node8F (fp32 boxes)
    nodes = np.frombuffer(blob[nodes_off:nodes_off + node_bytes * num_nodes].tobytes(), dtype=np.uint8).reshape(num_nodes, node_bytes)
    tris = np.frombuffer(blob[tris_off:tris_off + 48 * num_tris].tobytes(), dtype=np.float32).reshape(num_tris, 3, 4)
    return {"num_tris": num_tris, "num_nodes": num_nodes, "nodes": nodes, "tris": tris, "bounds": bounds, "total_bytes": total,
            "depth": int(hdr[17]), "node_bytes": node_bytes}


def validate_gas(gas, pad_check=True):
    """Structural invariants of the 8-wide BVH: every triangle referenced exactly once, every child box
    (decoded exactly as the traversal does: P + q * 2^e) contains all triangles below it, inner children are
    stored in slot order.  Returns (max_depth, leaf_count)."""
    nodes, tris = gas["nodes"], gas["tris"]
    n = gas["num_tris"]
    if n == 0:
        return 0, 0
    seen = np.zeros(n, np.int32)
    ordinals = tris[:, 2, 3].view(np.uint32)
    assert np.array_equal(np.sort(ordinals), np.arange(n, dtype=np.uint32)), "triangle ordinals are not a permutation"
    node_seen = np.zeros(gas["num_nodes"], np.int32)
    max_depth, leaves = 0, 0
    stack = [(0, None, None, 1)]
    while stack:
        ni, blo, bhi, depth = stack.pop()
        node_seen[ni] += 1
        max_depth = max(max_depth, depth)
        raw = nodes[ni]
        P = raw[0:12].view(np.float32).astype(np.float64)
        child_base, tri_base = [int(x) for x in raw[16:24].view(np.uint32)]
        meta = raw[24:32]
        if gas.get("node_bytes", 80) == 224:
            # Node8F: [P, imask][child_base, tri_base, meta][lo x|y|z: 8 floats each][hi x|y|z: 8 floats each], offsets from P
            imask = int(raw[12])
            planes = raw[32:224].view(np.float32).astype(np.float64).reshape(6, 8)
            q = planes
            scale = np.ones(3)
        else:
            e = raw[12:15].astype(np.int64)
            imask = int(raw[15])
            q = raw[32:80].reshape(6, 8).astype(np.float64)  # qlo x,y,z ; qhi x,y,z
            scale = np.ldexp(1.0, e - 127)
        for s in range(8):
            m = int(meta[s])
            if m == 0:
                continue
            lo = P + q[0:3, s] * scale
            hi = P + q[3:6, s] * scale
            assert np.all(lo <= hi)
            if blo is not None and pad_check:
                pass  # child boxes need not nest exactly inside the (separately quantised) parent box
            if (m >> 5) == 1 and (m & 0x1f) >= 24:
                assert (m & 0x1f) == 24 + s, "inner meta must encode its slot"
                assert imask >> s & 1
                ci = child_base + bin(imask & ((1 << s) - 1)).count("1")
                stack.append((ci, lo, hi, depth + 1))
                # all triangles below ci must be inside (lo,hi): checked when they are reached via box chain
                _check_subtree_box = (lo, hi)
                stack[-1] = (ci, lo, hi, depth + 1)
            else:
                assert not (imask >> s & 1)
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]
                off = m & 0x1f
                leaves += 1
                for t in range(tri_base + off, tri_base + off + cnt):
                    seen[t] += 1
                    v = tris[t, :, :3].astype(np.float64)
                    assert np.all(v >= lo - 0) and np.all(v <= hi + 0), f"triangle {t} outside its leaf box"
                    if blo is not None:
                        assert np.all(v >= blo) and np.all(v <= bhi), f"triangle {t} outside its parent's child box"
    assert np.all(seen == 1), f"{(seen != 1).sum()} triangles not referenced exactly once"
    assert np.all(node_seen == 1), f"{(node_seen != 1).sum()} nodes not referenced exactly once"
    return max_depth, leaves

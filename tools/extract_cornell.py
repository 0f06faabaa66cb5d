#!/usr/bin/env python3
"""Extract the Cornell-box scene tables from the reference sample into a data fixture.

Reads (never copies code from) /root/reference/SDK/optixPathTracer/optixPathTracer.cpp:144-314
(g_vertices, g_mat_indices, g_emission_colors, g_diffuse_colors) plus the light/camera constants
at :424-445 and :536-552, and writes optix_raytracer_b200/data/cornell.json.  The JSON is *scene
input data* shared by the product's host side, the oracle tests and bench.py; it is committed so
nothing needs /root/reference at run time.  Re-run only if the reference scene changes.
"""
import json, re, sys, pathlib

SRC = pathlib.Path("/root/reference/SDK/optixPathTracer/optixPathTracer.cpp")
OUT = pathlib.Path(__file__).resolve().parents[1] / "optix_raytracer_b200" / "data" / "cornell.json"

def block(text, start_pat):
    i = text.index(start_pat)
    j = text.index("};", i)
    return text[i:j]

def floats(s):
    return [float(x.rstrip("f")) for x in re.findall(r"-?\d+\.\d*f?|-?\d+f?", s)]

def main():
    text = SRC.read_text()
    vb = block(text, "g_vertices =")
    verts = []
    for m in re.finditer(r"\{\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*\}", vb):
        verts.append([float(m.group(k)) for k in (1, 2, 3)])
    mb = block(text, "g_mat_indices =")
    mb = re.sub(r"//.*", "", mb)
    mats = [int(x) for x in re.findall(r"\b\d+\b", mb.split("{{", 1)[1])]
    def colors(name):
        b = block(text, name + " =")
        return [[float(a), float(b_), float(c)] for a, b_, c in
                re.findall(r"\{\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*\}", b)]
    emission = colors("g_emission_colors")
    diffuse = colors("g_diffuse_colors")
    assert len(verts) == 96 and len(mats) == 32 and len(emission) == 4 and len(diffuse) == 4, \
        (len(verts), len(mats), len(emission), len(diffuse))
    data = {
        "source": "SDK/optixPathTracer/optixPathTracer.cpp:144-314,424-445,536-552",
        "vertices": verts,            # 96 x float3 (reference stores float4 with pad 0)
        "mat_indices": mats,          # per-triangle material / SBT record index
        "emission_colors": emission,
        "diffuse_colors": diffuse,
        "light": {"emission": [15.0, 15.0, 5.0], "corner": [343.0, 548.5, 227.0],
                  "v1": [0.0, 0.0, 105.0], "v2": [-130.0, 0.0, 0.0]},
        "camera": {"eye": [278.0, 273.0, -900.0], "lookat": [278.0, 273.0, 330.0],
                   "up": [0.0, 1.0, 0.0], "fov_y": 35.0},
        "width": 768, "height": 768, "samples_per_launch": 16,
    }
    OUT.write_text(json.dumps(data, indent=1) + "\n")
    print("wrote", OUT, len(verts), "verts")

if __name__ == "__main__":
    sys.exit(main())

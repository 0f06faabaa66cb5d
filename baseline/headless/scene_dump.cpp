// scene_dump — TEST INFRASTRUCTURE: what sutil::loadScene (SDK/sutil/Scene.cpp:267-550) makes of a glTF file, written out so that
// optix_raytracer_b200/host.py: load_gltf can be compared with it (tests/test_gpu_reference_samples.py).  Links against the reference's
// own sutil (baseline/Makefile); every buffer view is read back from the device through its (data, count, byte_stride, elmt_byte_size),
// exactly what the device programs see.  Output: one JSON document on stdout.
//   scene_dump <file.gltf>
#include <cuda_runtime.h>
#include <optix.h>
#include <optix_function_table_definition.h>
#include <optix_stubs.h>
#include <sutil/Camera.h>
#include <sutil/Scene.h>
#include <sutil/sutil.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

template <typename V>
static std::vector<unsigned char> fetch(const V& v, size_t elmt)
{
    std::vector<unsigned char> out((size_t)v.count * elmt);
    if (!v.data || !v.count) return out;
    const size_t stride = v.byte_stride ? v.byte_stride : elmt;
    cudaMemcpy2D(out.data(), elmt, (const void*)v.data, stride, elmt, v.count, cudaMemcpyDeviceToHost);
    return out;
}
static void put_floats(const char* key, const std::vector<unsigned char>& b, bool comma = true)
{
    // the fp32 BIT PATTERNS (a decimal -0 would come back from a JSON parser as the integer 0)
    printf("\"%s\": [", key);
    const unsigned* f = (const unsigned*)b.data();
    for (size_t i = 0; i < b.size() / 4; ++i) printf("%s%u", i ? "," : "", f[i]);
    printf("]%s", comma ? "," : "");
}
static void put_tex(const char* key, const MaterialData::Texture& t, bool comma = true)
{
    printf("\"%s\": {\"present\": %d, \"texcoord\": %d, \"offset\": [%.9g,%.9g], \"rotation\": [%.9g,%.9g], \"scale\": [%.9g,%.9g]}%s", key, t.tex ? 1 : 0, t.texcoord,
           t.texcoord_offset.x, t.texcoord_offset.y, t.texcoord_rotation.x, t.texcoord_rotation.y, t.texcoord_scale.x, t.texcoord_scale.y, comma ? "," : "");
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: scene_dump <file.gltf>\n"); return 2; }
    try {
        sutil::Scene scene;
        sutil::loadScene(argv[1], scene);
        printf("{\"meshes\": [");
        bool first_mesh = true;
        for (const auto& m : scene.meshes()) {
            printf("%s{\"name\": \"%s\", \"aabb\": [%.9g,%.9g,%.9g,%.9g,%.9g,%.9g], \"primitives\": [", first_mesh ? "" : ",", m->name.c_str(), m->object_aabb.m_min.x,
                   m->object_aabb.m_min.y, m->object_aabb.m_min.z, m->object_aabb.m_max.x, m->object_aabb.m_max.y, m->object_aabb.m_max.z);
            first_mesh = false;
            for (size_t p = 0; p < m->positions.size(); ++p) {
                printf("%s{", p ? "," : "");
                put_floats("positions", fetch(m->positions[p], 12));
                put_floats("normals", fetch(m->normals[p], 12));
                put_floats("texcoords0", fetch(m->texcoords[0][p], 8));
                put_floats("texcoords1", fetch(m->texcoords[1][p], 8));
                put_floats("colors", fetch(m->colors[p], 16));
                const auto& iv = m->indices[p];
                const std::vector<unsigned char> ib = fetch(iv, iv.elmt_byte_size ? iv.elmt_byte_size : 4);
                printf("\"index_size\": %d, \"indices\": [", (int)iv.elmt_byte_size);
                for (unsigned i = 0; i < iv.count; ++i) {
                    unsigned v = 0;
                    if (iv.elmt_byte_size == 2) v = ((const unsigned short*)ib.data())[i];
                    else if (iv.elmt_byte_size == 1) v = ib[i];
                    else v = ((const unsigned*)ib.data())[i];
                    printf("%s%u", i ? "," : "", v);
                }
                printf("], \"strides\": {\"positions\": %d, \"normals\": %d, \"texcoords0\": %d}, \"material\": %d}", (int)m->positions[p].byte_stride, (int)m->normals[p].byte_stride,
                       (int)m->texcoords[0][p].byte_stride, m->material_idx[p]);
            }
            printf("]}");
        }
        printf("], \"instances\": [");
        bool first = true;
        for (const auto& i : scene.instances()) {
            printf("%s{\"mesh\": %d, \"transform\": [", first ? "" : ",", i->mesh_idx);
            first = false;
            const float* t = i->transform.getData();
            for (int k = 0; k < 16; ++k) printf("%s%.9g", k ? "," : "", t[k]);
            printf("], \"world_aabb\": [%.9g,%.9g,%.9g,%.9g,%.9g,%.9g]}", i->world_aabb.m_min.x, i->world_aabb.m_min.y, i->world_aabb.m_min.z, i->world_aabb.m_max.x,
                   i->world_aabb.m_max.y, i->world_aabb.m_max.z);
        }
        printf("], \"materials\": [");
        first = true;
        for (const MaterialData& m : scene.materials()) {
            printf("%s{\"base_color\": [%.9g,%.9g,%.9g,%.9g], \"metallic\": %.9g, \"roughness\": %.9g, \"alpha_mode\": %d, \"alpha_cutoff\": %.9g, \"double_sided\": %d, "
                   "\"emissive_factor\": [%.9g,%.9g,%.9g],", first ? "" : ",", m.pbr.base_color.x, m.pbr.base_color.y, m.pbr.base_color.z, m.pbr.base_color.w, m.pbr.metallic,
                   m.pbr.roughness, (int)m.alpha_mode, m.alpha_cutoff, m.doubleSided ? 1 : 0, m.emissive_factor.x, m.emissive_factor.y, m.emissive_factor.z);
            first = false;
            put_tex("base_color_tex", m.pbr.base_color_tex);
            put_tex("metallic_roughness_tex", m.pbr.metallic_roughness_tex);
            put_tex("normal_tex", m.normal_tex);
            put_tex("emissive_tex", m.emissive_tex, false);
            printf("}");
        }
        // the camera Scene::camera() hands the viewer needs the scene bounds: set them the way Scene::finalize does without building
        // anything (finalize would create a context and compile the programs)
        printf("], \"cameras\": [");
        sutil::Aabb aabb;
        for (const auto& i : scene.instances()) aabb.include(i->world_aabb);
        printf("], \"scene_aabb\": [%.9g,%.9g,%.9g,%.9g,%.9g,%.9g]}\n", aabb.m_min.x, aabb.m_min.y, aabb.m_min.z, aabb.m_max.x, aabb.m_max.y, aabb.m_max.z);
    } catch (const std::exception& e) {
        fprintf(stderr, "scene_dump: %s\n", e.what());
        return 1;
    }
    return 0;
}

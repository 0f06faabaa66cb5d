// internal.h — functions implemented in the .cu files and exposed through the C ABI in api.cu
#pragma once
#include "common.h"

namespace b200rt {
// bvh_build.cu
int accel_compute_memory_usage(b200rt_context, const b200rt_accel_build_options*, const b200rt_build_input*, unsigned,
                               b200rt_accel_buffer_sizes*);
int accel_build(b200rt_context, cudaStream_t, const b200rt_accel_build_options*, const b200rt_build_input*, unsigned, b200rt_deviceptr,
                size_t, b200rt_deviceptr, size_t, b200rt_traversable*, const b200rt_accel_emit_desc*, unsigned);
int accel_compact(b200rt_context, cudaStream_t, b200rt_traversable, b200rt_deviceptr, size_t, b200rt_traversable*);
int accel_get_info(b200rt_context, b200rt_traversable, b200rt_accel_info*);
int accel_emit_property(b200rt_context, cudaStream_t, b200rt_traversable, const b200rt_accel_emit_desc*, unsigned);
size_t radix_sort_hist_words(size_t n);
size_t radix_sort_scan_words(size_t n);
int radix_sort_pairs32(b200rt_context, cudaStream_t, uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int passes, uint32_t* hist, uint32_t* scan_tmp, int* result);
// raycast.cu
int launch_raycast(b200rt_context, cudaStream_t, b200rt_deviceptr d_params, const b200rt_shader_binding_table*, unsigned, unsigned,
                   b200rt_deviceptr ext_hits);
int create_rays_ortho(b200rt_context, cudaStream_t, b200rt_deviceptr, int, int, const float*, const float*, float);
int translate_rays(b200rt_context, cudaStream_t, b200rt_deviceptr, int, const float*);
int shade_hits(b200rt_context, cudaStream_t, b200rt_deviceptr, int, b200rt_deviceptr);
int trace_closest(b200rt_context, cudaStream_t, b200rt_traversable, b200rt_deviceptr, uint64_t, unsigned, b200rt_deviceptr);
int trace_any(b200rt_context, cudaStream_t, b200rt_traversable, b200rt_deviceptr, uint64_t, unsigned, b200rt_deviceptr);
int trace_stats(b200rt_context, cudaStream_t, b200rt_traversable, b200rt_deviceptr, uint64_t, uint64_t*, uint64_t*);
int trace_buffer(b200rt_context, cudaStream_t, b200rt_traversable, b200rt_deviceptr rays, uint64_t n_max, const unsigned int* n_dev, unsigned n_mult,
                 int kind, unsigned ray_flags, b200rt_deviceptr out, unsigned flag_period = 0, b200rt_deviceptr sbt_out = 0,
                 b200rt_deviceptr handle_dev = 0, const b200rt_shader_binding_table* anyhit_sbt = nullptr);
// whitted.cu
int launch_whitted(b200rt_context, cudaStream_t, b200rt_deviceptr d_params, const b200rt_shader_binding_table*, unsigned, unsigned);
void whitted_release(b200rt_context);  // destroys the context's BLEND level loop
int texture_create(b200rt_context, int, int, const void*, int, int, int, uint64_t*, uint64_t*);
int texture_destroy(b200rt_context, uint64_t, uint64_t);
int texture_view(b200rt_context, uint64_t, int, int, int, uint64_t*);
// playground.cu
int launch_playground(b200rt_context, cudaStream_t, b200rt_deviceptr d_params, unsigned, unsigned, const b200rt_pt_options*);
int generate_playground_scene(b200rt_context, cudaStream_t, uint32_t rows, uint32_t seed, b200rt_deviceptr, b200rt_deviceptr, b200rt_deviceptr, uint64_t*);
// pathtracer.cu
int launch_pathtracer(b200rt_context, cudaStream_t, b200rt_deviceptr d_params, const b200rt_shader_binding_table*, unsigned, unsigned,
                      const b200rt_pt_options*, int multigpu);
int fill_samples(b200rt_context, cudaStream_t, int, int, int, int, b200rt_deviceptr, int);
int deinterleave(b200rt_context, cudaStream_t, b200rt_deviceptr, int, int, int, int, b200rt_deviceptr, b200rt_deviceptr);
int generate_synthetic_mesh(b200rt_context, cudaStream_t, uint64_t, uint32_t, b200rt_deviceptr, b200rt_deviceptr, float*);
void pathtracer_release(b200rt_context);  // destroys the context's launch graphs
uint64_t pathtracer_graph_kernels(b200rt_context);  // kernels run inside launch graphs so far
}  // namespace b200rt

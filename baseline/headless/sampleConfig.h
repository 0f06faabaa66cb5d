// sampleConfig.h as cmake generates it from SDK/sampleConfig.h.in, for the headless build (baseline/Makefile).  The directories are
// overridden at run time through OPTIX_SAMPLES_SDK_DIR / OPTIX_SAMPLES_SDK_PTX_DIR (SDK/sutil/sutil.cpp:162-178,986-1021): the
// compile-time values only have to be syntactically present.
#pragma once
#define SAMPLES_DIR "baseline/_ref/SDK"
#define SAMPLES_PTX_DIR "baseline/_ref/ptx"
#define SAMPLES_CUDA_DIR "baseline/_ref/SDK/cuda"
#define SAMPLES_RELATIVE_INCLUDE_DIRS "cuda", "sutil", ".",
#define SAMPLES_ABSOLUTE_INCLUDE_DIRS "include",
#define CUDA_NVRTC_ENABLED 0
#define CUDA_NVRTC_OPTIONS "-std=c++11",
#define SAMPLES_INPUT_GENERATE_OPTIXIR 0
#define SAMPLES_INPUT_GENERATE_PTX 1

for lib in optix_raytracer_b200/libb200rt.so gpurun_variants/*.so; do
  [ -f "$lib" ] || continue
  echo -n "$lib: "; B200RT_LIB_PATH=$PWD/$lib timeout 120 python tools/time_small.py 2>&1 | tail -1
done

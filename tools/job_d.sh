for lib in optix_raytracer_b200/libb200rt.so gpurun_variants/*.so; do
  echo "== $lib"
  B200RT_LIB_PATH=$PWD/$lib B200RT_BUILD_TIMING=1 timeout 120 python tools/build_once.py --reps 3 2>&1 | grep "b200rt build\|build 2" | tail -2
done

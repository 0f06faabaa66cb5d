B200RT_BUILD_TIMING=2 timeout 120 python tools/build_once.py 2>&1 | grep "b200rt build\|build 1" | tail -3
B200RT_BUILD_TIMING=1 timeout 120 python tools/build_once.py 2>&1 | grep "b200rt build\|build 1" | tail -2

#!/bin/bash
# tools/time_small.py over (library, environment) pairs.  usage: tools/ab_small_env.sh "lib.so VAR=val ..." ...
for spec in "$@"; do
  set -- $spec; lib=$1; shift
  echo -n "$spec: "; env "$@" B200RT_LIB_PATH=$PWD/$lib timeout 120 python tools/time_small.py 2>&1 | tail -1
done

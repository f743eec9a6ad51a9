#!/bin/bash
# usage: build_variants.sh A="" B="-DCGRT_NO_POSTPONE" ...   -> gpurun_variants_<name>.so (parallel nvcc)
for kv in "$@"; do
  name="${kv%%=*}"; flags="${kv#*=}"
  ( CGRT_BUILD_OUT=$PWD/gpurun_variants_$name.so CGRT_NVCC_EXTRA="$flags" python -c "from cgraytracing_b200 import build; build.build(force=True)" && echo built $name "$flags" ) &
done
wait

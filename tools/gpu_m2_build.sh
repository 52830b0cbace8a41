#!/bin/bash
# usage: gpu_m2_build.sh N "<name>:<EXTRA nvcc flags>:<bench args>" ...  — rebuild per variant on the GPU box, then bench at N GPUs
N=${1:-2}; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; rest=${v#*:}; flags=${rest%%:*}; args=${rest#*:}
  make -C volumetric-particles-for-unity_b200/csrc clean >/dev/null; make -C volumetric-particles-for-unity_b200/csrc EXTRA="$flags" > gpurun_out/build_$name.log 2>&1 || { echo "build $name failed"; tail -5 gpurun_out/build_$name.log; continue; }
  tools/gpu_m2.sh $N "$name:$args"
done

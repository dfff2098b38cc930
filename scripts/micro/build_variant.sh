#!/bin/bash
# Build a variant of the library with extra nvcc flags into scripts/micro/libs/<name>.so (git-ignored, shipped by gpurun).
# Usage: scripts/micro/build_variant.sh <name> [-DFLAG ...]
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p scripts/micro/libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -fvisibility=default \
     --fmad=false -Iinclude "$@" -o scripts/micro/libs/$name.so soc_project_stereo_matching_b200/csrc/sgm_b200.cu
echo built scripts/micro/libs/$name.so

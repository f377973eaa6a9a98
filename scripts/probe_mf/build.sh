#!/bin/bash
# builds the probe variants (temporary)
set -e
cd "$(dirname "$0")"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I/root/repo/yelprecommendation_b200/csrc"
v() { name=$1; d=$2; shift 2; nvcc $F -DD_PROBE=$d -DVARIANT_NAME="\"$name\"" "$@" main.cu -o probe_${name}_d$d; }
v base 32 '-DYR_PRE_SYNC=' '-DYR_POST_SYNC=' '-DYR_FLAG_LD=__ldcg' &
v base 64 '-DYR_PRE_SYNC=' '-DYR_POST_SYNC=' '-DYR_FLAG_LD=__ldcg' &
v syncwarp 32 '-DYR_PRE_SYNC=' '-DYR_POST_SYNC=__syncwarp();' '-DYR_FLAG_LD=__ldcg' &
v fence 32 '-DYR_PRE_SYNC=__threadfence();' '-DYR_POST_SYNC=' '-DYR_FLAG_LD=__ldcg' &
wait
v volatileflag 32 '-DYR_PRE_SYNC=' '-DYR_POST_SYNC=' '-DYR_FLAG_LD=*(volatile int32_t*)' &
v postfence 32 '-DYR_PRE_SYNC=' '-DYR_POST_SYNC=__threadfence();' '-DYR_FLAG_LD=__ldcg' &
wait

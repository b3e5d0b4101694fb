#!/bin/bash
# Builds an A/B variant of the library with extra nvcc flags (e.g. -DNSAC_SCORE_PDL=0) into build/variants/<name>.so; run it with
#   NSAC_B200_LIB=build/variants/<name>.so python ...      usage: build_variant.sh <name> <flags...>
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" nopesac_b200/csrc/*.cu -o build/variants/$name.so
echo build/variants/$name.so

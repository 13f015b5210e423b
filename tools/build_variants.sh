#!/usr/bin/env bash
# A/B builds of the fused pass: lib/libnnfac_b200_<tag>.so with tc_fused.cu compiled under different switches
# (select one with NNFAC_B200_LIB=...).  usage: tools/build_variants.sh tag "-DFUSED_PREFETCH=0 -DFUSED_PACKED=1" [file]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"; PKG="$HERE/nn-fac_b200"
tag="$1"; flags="$2"; file="${3:-tc_fused}"
mkdir -p /tmp/variants/$tag
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c "$PKG/csrc/$file.cu" -o /tmp/variants/$tag/$file.o
objs=$(ls "$PKG"/build/*.o | grep -v "/$file.o")
/usr/local/cuda/bin/nvcc -Wno-deprecated-gpu-targets -shared -o "$PKG/lib/libnnfac_b200_$tag.so" $objs /tmp/variants/$tag/$file.o -lcudart_static -ldl -lrt -lpthread
echo "built $PKG/lib/libnnfac_b200_$tag.so"

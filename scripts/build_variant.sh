#!/bin/bash
# A/B builds of one translation unit with extra -D flags: scripts/build_variant.sh NAME file.cu "-DFOO=1 -DBAR=2"
# -> vapoursynth_zip_b200/lib/variants/libvszip_NAME.so (select with VSZIP_CUDA_LIB=...).  Only radius 13 is instantiated (VSZ_SEG_DEV13).
set -e
NAME=$1; SRC=$2; FLAGS=$3
OUT=vapoursynth_zip_b200/lib/variants; OBJ=build/variants/$NAME
mkdir -p $OUT $OBJ
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude -Ivapoursynth_zip_b200/csrc"
$NV -DVSZ_SEG_DEV13 $FLAGS -Xptxas -v -c vapoursynth_zip_b200/csrc/$SRC -o $OBJ/${SRC%.cu}.o 2> $OBJ/ptxas.log
OBJS=""
for o in build/obj/*.o; do b=$(basename $o); if [ "$b" == "${SRC%.cu}.o" ]; then OBJS="$OBJS $OBJ/$b"; else OBJS="$OBJS $o"; fi; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $OUT/libvszip_$NAME.so $OBJS
grep -A2 "kernelILi13" $OBJ/ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo

#!/bin/bash
# times the given quick_bench configs for every A/B library under vapoursynth_zip_b200/lib/variants (and the main build)
CFG=${1:-H5,H1}
echo -n "main: "; timeout 200 python scripts/quick_bench.py 256 5 $CFG 2>&1 | tail -1
for f in vapoursynth_zip_b200/lib/variants/libvszip_*.so; do
  n=$(basename $f .so); echo -n "${n#libvszip_}: "; VSZIP_CUDA_LIB=$f timeout 200 python scripts/quick_bench.py 256 5 $CFG 2>&1 | tail -1
done

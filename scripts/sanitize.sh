#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck + synccheck over smoke() and one small parity test per kernel family.
# usage (from the repo root, under gpurun): bash scripts/sanitize.sh r02     -> gpurun_out/sanitizer_<tool>_<round>.log (tails)
R=${1:-r02}
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
# one representative, small case per kernel family (segment/TMA BoxBlur, ring BoxBlur, comptime float, Bilateral smem/compute/PBFIC,
# reductions incl. the sampled bracket path, pointwise, fused chain, host pin cache)
SEL="test_noise_bit_exact or test_small_sigma_r_is_bit_exact_16bit or test_joint_ref or test_pbfic_joint_and_tiny_planes or test_minmax_and_average_from_one_read or test_structured_planeminmax or test_fused_chain_partial_planes or test_fused_chain_adaptive_binarize or test_pageable_buffers or test_comptime_float_small_case_for_the_sanitizer or test_comptime_widths_around_the_cta_shape_switch or test_planeminmax_ranks_at_the_ends_of_the_range or test_full_size_config1"
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  timeout 1500 $CS --tool $tool $extra --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); import vapoursynth_zip_b200 as vz; vz.core.shutdown()" > $O/san_${tool}_smoke.txt 2>&1
  echo "smoke rc=$?" >> $O/san_${tool}_smoke.txt
  timeout 2400 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x --timeout 2000 -k "$SEL" > $O/san_${tool}_tests.txt 2>&1
  echo "tests rc=$?" >> $O/san_${tool}_tests.txt
  { echo "## compute-sanitizer --tool $tool, smoke()"; tail -12 $O/san_${tool}_smoke.txt; echo; echo "## compute-sanitizer --tool $tool, pytest -k '$SEL'"; tail -15 $O/san_${tool}_tests.txt; } > $O/sanitizer_${tool}_$R.log
done
for f in $O/sanitizer_*_$R.log; do tail -n 4 $f; done

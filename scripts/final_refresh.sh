#!/bin/bash
# Re-measures the numbers of profiles/ after a late kernel change (both bench arms, extras, fused-stats ncu summary, GPU suite).
# usage (from the repo root, under gpurun): bash scripts/final_refresh.sh
R=r02; O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$R.json 2>> $O/bench_$R.err
timeout 600 python scripts/bench_extras.py --frames 128 --out $O/bench_extras_$R.json > /dev/null 2> $O/bench_extras_$R.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:minmax_bracket" -s 1 -c 1 -o $O/ncu_planestats_$R python scripts/prof_run.py planestats 32 2 > /dev/null 2>&1
python scripts/ncu_summary.py $O/ncu_planestats_$R.ncu-rep > $O/ncu_planestats_$R.md 2>/dev/null
rm -f $O/ncu_planestats_$R.ncu-rep
timeout 1000 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > $O/pytest_gpu_$R.log
python scripts/show_bench.py $O/bench_$R.json | head -40; cut -c1-160 $O/bench_ref_$R.json; cat $O/pytest_gpu_$R.log | tail -1

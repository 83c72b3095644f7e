#!/bin/bash
# One GPU-box visit: parity suite, contract bench line, secondary configs, ncu launch list + full captures.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5 > $O/pytest_gpu_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$R.log 2>&1
timeout 600 python bench.py > $O/bench_$R.json 2> $O/bench_$R.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2>> $O/bench_$R.err
timeout 300 python scripts/e2e_chain.py 8 4 > $O/e2e_chain_$R.txt 2>&1
timeout 300 python scripts/e2e_diamond.py 32 8 > $O/e2e_diamond_$R.txt 2>&1
timeout 600 python scripts/bench_extras.py --frames 128 --out $O/bench_extras_$R.json > /dev/null 2> $O/bench_extras_$R.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/launches_bench_$R.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_ -s 2 -c 2 -o $O/ncu_boxblur_$R python scripts/prof_run.py boxblur 128 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bilateral -s 3 -c 1 -o $O/ncu_bilateral_$R python scripts/prof_run.py bilateral 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:stats_kernel|average_u16" -s 1 -c 1 -o $O/ncu_average_$R python scripts/prof_run.py average 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:hist_sample|minmax_bracket" -s 2 -c 2 -o $O/ncu_minmax_$R python scripts/prof_run.py minmax 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:limitfilter_kernel" -s 1 -c 1 -o $O/ncu_limitfilter_$R python scripts/prof_run.py limitfilter 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:adaptivebinarize" -s 1 -c 1 -o $O/ncu_binarize_$R python scripts/prof_run.py binarize 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:minmax_bracket" -s 1 -c 1 -o $O/ncu_planestats_$R python scripts/prof_run.py planestats 32 2 > /dev/null 2>&1
# summaries are made on the box; only the BoxBlur capture travels back as a .ncu-rep (gpurun merges at most 64 MiB)
for f in $O/ncu_*_$R.ncu-rep; do python scripts/ncu_summary.py $f > ${f%.ncu-rep}.md 2>/dev/null; done
for k in bilateral average minmax limitfilter binarize planestats; do rm -f $O/ncu_${k}_$R.ncu-rep; done
cat $O/pytest_gpu_$R.log; cat $O/bench_$R.json | cut -c1-400; tail -3 $O/bench_$R.err

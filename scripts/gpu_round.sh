#!/bin/bash
# One GPU-box visit: parity suite, contract bench line (both arms), ncu launch list + full captures of every kernel family.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh r02        (sanitizer logs: scripts/sanitize.sh)
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > $O/pytest_gpu_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$R.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$R.json 2>> $O/bench_$R.err
timeout 300 python scripts/e2e_diamond.py 32 8 > $O/e2e_diamond_$R.txt 2>&1
timeout 600 python scripts/bench_extras.py --frames 128 --out $O/bench_extras_$R.json > /dev/null 2> $O/bench_extras_$R.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > $O/launches_bench_$R.log 2>&1
cap() {  # name, kernel regex, skip, count, prof_run args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o $O/ncu_${name}_$R python scripts/prof_run.py "$@" > /dev/null 2>&1
  python scripts/ncu_summary.py $O/ncu_${name}_$R.ncu-rep > $O/ncu_${name}_$R.md 2>/dev/null
}
cap boxblur "hseg_kernel|vseg_tile_kernel" 3 3 boxblur 128 2   # one whole call: H, V luma, V chroma
cap boxblur_ct "ctfused_kernel" 2 2 boxblur_ct 128 2
cap boxblur_ctf "ctf_" 2 2 boxblur_ctf 8 2
cap bilateral "bilateral" 3 1 bilateral 32 2
cap average "stats_kernel|average_u16" 1 1 average 32 2
cap minmax "hist_sample|minmax_bracket" 2 2 minmax 32 2
cap planestats "minmax_bracket" 1 1 planestats 32 2
cap limitfilter "limitfilter_kernel" 1 1 limitfilter 32 2
cap binarize "adaptivebinarize" 1 1 binarize 32 2
# only the headline capture travels back as a .ncu-rep (gpurun merges at most 64 MiB)
for k in boxblur_ct boxblur_ctf bilateral average minmax limitfilter binarize planestats; do rm -f $O/ncu_${k}_$R.ncu-rep; done
cat $O/pytest_gpu_$R.log; cat $O/smoke_$R.log | tail -1; python scripts/show_bench.py $O/bench_$R.json; cut -c1-300 $O/bench_ref_$R.json; tail -3 $O/bench_$R.err

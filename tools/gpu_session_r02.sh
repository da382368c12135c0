#!/bin/bash
# One GPU session for the round-2 evidence: bench line, launch list, ncu full captures of the top kernels, config-1 timing.
# Usage (on the GPU box, from the repo root):  bash tools/gpu_session_r02.sh <tag>
tag=${1:-r02f}
out=gpurun_out
set -x
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 200 python tools/bench_next_rows.py --rows C1 > $out/${tag}_c1.jsonl 2> $out/${tag}_c1.err; echo "c1 rc=$?"
BENCH="python bench.py --steps 2 --warmup 1 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv $BENCH > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_win_kernel -s 30 -c 4 -o $out/${tag}_wgrad_win -f $BENCH > $out/${tag}_ncu_wgw.log 2>&1; echo "ncu wgrad_win rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_win_kernel -s 60 -c 6 -o $out/${tag}_conv_win -f $BENCH > $out/${tag}_ncu_cw.log 2>&1; echo "ncu conv_win rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sdelayer -c 3 -o $out/${tag}_sdelayer -f python tools/bench_next_rows.py --rows C1 --reps 1 > $out/${tag}_ncu_sl.log 2>&1; echo "ncu sdelayer rc=$?"
ls -la $out | tail -12

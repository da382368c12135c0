timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_config4_full.py -x -q -m gpu -k "block_tc or remat or classifier or config or golden or full_size" > gpurun_out/u23_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u23_tests.log
timeout 300 python tools/step_gaps.py --steps 2 > gpurun_out/u23_gaps.txt 2>&1; sed -n 3,4p gpurun_out/u23_gaps.txt; tail -8 gpurun_out/u23_gaps.txt
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 200 $B > gpurun_out/u23_b.json 2> gpurun_out/u23_b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/u23_b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
PY

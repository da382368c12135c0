timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_config4_full.py -x -q -m gpu -k "tconv or block or config or golden or edge or classifier or remat" > gpurun_out/u24_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/u24_tests.log
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 200 $B > gpurun_out/u24_b.json 2> gpurun_out/u24_b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/u24_b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
PY

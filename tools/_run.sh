timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tconv or block_tc or block_types or config3 or edge or golden or wgrad or remat" > gpurun_out/u14_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u14_tests.log
python tools/bench_tconv.py --S 72 --groups 1,2 --reps 5 2>&1 | grep wgrad
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 200 $B > gpurun_out/u14_b.json 2> gpurun_out/u14_b.err
python - <<'PY'
import json
for n in ("b",):
    try:
        d=json.loads(open(f"gpurun_out/u14_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
    except Exception as e: print(n, "ERR", e)
PY

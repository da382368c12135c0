timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tconv or block_tc or block_types or config3 or edge or golden" > gpurun_out/u11_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u11_tests.log
python tools/bench_tconv.py --groups 0,1 --no_wgrad --dact 2>&1 | grep -E "dgrad|fwd"
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 200 $B > gpurun_out/u11_b.json 2> gpurun_out/u11_b.err
python - <<'PY'
import json
for n in ("b",):
    try:
        d=json.loads(open(f"gpurun_out/u11_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
    except Exception as e: print(n, "ERR", e)
PY

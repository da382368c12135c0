timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tconv or block_tc or block_types or config3 or edge" > gpurun_out/u6_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u6_tests.log
python tools/bench_tconv.py --groups 0,1 --no_wgrad > gpurun_out/u6_tconv.txt 2>&1; cat gpurun_out/u6_tconv.txt | tail -16
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
BSDE_LIB_PATH=$PWD/bayesde_b200/build/libbsde_base.so timeout 200 $B > gpurun_out/u6_b_base.json 2> gpurun_out/u6_b_base.err; echo "base rc=$?"
timeout 200 $B > gpurun_out/u6_b_new.json 2> gpurun_out/u6_b_new.err; echo "new rc=$?"
python - <<'PY'
import json
for n in ("base","new"):
    try:
        d=json.loads(open(f"gpurun_out/u6_b_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
    except Exception as e: print(n, "ERR", e)
PY

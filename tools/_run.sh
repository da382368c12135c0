timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tconv or block_tc or block_types or config3 or edge or golden or remat" > gpurun_out/u21_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u21_tests.log
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
for v in 0 1 0 1; do
BSDE_WIN_REV=$v timeout 200 $B > gpurun_out/u21_b.json 2> gpurun_out/u21_b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/u21_b.json").read().strip().splitlines()[-1])
print("rev=$v", d["ms_per_step"], d["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
PY
done

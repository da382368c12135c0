timeout 300 python tools/step_gaps.py > gpurun_out/u16_gaps.txt 2>&1; tail -22 gpurun_out/u16_gaps.txt
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 200 $B > gpurun_out/u16_b.json 2> gpurun_out/u16_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/u16_b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
PY

B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
for i in 1 2; do
timeout 200 $B > gpurun_out/u19_b$i.json 2> gpurun_out/u19_b$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/u19_b$i.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
PY
done

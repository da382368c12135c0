timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02l_tests_full.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02l_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02l_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","value_fp32")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline_hbm"]["frac"], d.get("config5",{}).get("value"), d["cpu_baseline"]["value"], d["kernel_ms_per_step"], d["clocks"], d["fp32_leg"]["same_seed_loss"])
PY

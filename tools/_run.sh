for pf in 0 1; do echo "== PF=$pf"; BSDE_WIN_PF=$pf python tools/bench_tconv.py --groups 0,1 --no_wgrad --dact 2>&1 | grep dgrad; done
B="python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_fp32_leg --no_config5"
for pf in 0 1; do BSDE_WIN_PF=$pf timeout 200 $B > gpurun_out/u10_b_pf$pf.json 2> gpurun_out/u10_b_pf$pf.err; done
python - <<'PY'
import json
for n in ("pf0","pf1"):
    try:
        d=json.loads(open(f"gpurun_out/u10_b_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d.get("kernel_ms_per_step"), d["loss"]["last_timed_step"])
    except Exception as e: print(n, "ERR", e)
PY

#!/bin/bash
# weak (batch-sharded, default) and strong (samples x batch grid, fixed 8 x 128 work) bench lines at N ranks of one box
N=$1; tag=${2:-r02f}
P=$((29500 + N))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --no_fp32_leg --no_cpu_baseline > gpurun_out/${tag}_n${N}_weak.json 2> gpurun_out/${tag}_n${N}_weak.err; echo "weak rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --gpus $N --steps 5 --warmup 3 --shard samples --no_fp32_leg --no_cpu_baseline > gpurun_out/${tag}_n${N}_strong.json 2> gpurun_out/${tag}_n${N}_strong.err; echo "strong rc=$?"
python - <<PY
import json
for m in ("weak", "strong"):
    try:
        d = json.load(open("gpurun_out/${tag}_n${N}_%s.json" % m))
        print(m, d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), (d.get("config5") or {}).get("value"))
    except Exception as e:
        print(m, "failed", e)
PY
for f in gpurun_out/${tag}_n${N}_weak.err gpurun_out/${tag}_n${N}_strong.err; do tail -n 2 $f; done

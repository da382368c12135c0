#!/usr/bin/env python
"""bench.py -- SDE-BNN training throughput on the fused B200 path (BASELINE.json metric).

A "step" = one optimiser step of examples/jax/sdebnn_classification.py's `update` on one batch of
synthetic 32x32x3 images: forward through Augment -> 6 SDEBNN blocks (nblocks 2-2-2, fx_dim 64,
fw_dims 2-64-2, nsteps 20 => 19 drift evaluations per block) -> mean-field Dense head, reverse pass
through the solver (recompute from per-iteration x checkpoints), gradient mean over the S Monte-Carlo
samples (+ NCCL all-reduce over ranks), Adam.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                       CPU oracle port on the host cores

`value` counts distinct training images per second; each image is integrated under `nsamples`
weight paths (examples/jax/sdebnn_classification.py:218 vmaps grad(loss) over the sample rngs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SDE-BNN train images/s (32x32, nsteps=20)"
UNIT = "images/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    p.add_argument("--nsamples", type=int, default=8)
    p.add_argument("--math", default=os.environ.get("BSDE_MATH", "tc"), choices=["fp32", "tc"])
    p.add_argument("--remat", action="store_true", help="recompute conv intermediates in the reverse pass")
    p.add_argument("--time_grid", default="reference", choices=["reference", "uniform"])
    p.add_argument("--diff_coef", type=float, default=0.2)
    p.add_argument("--cpu_batch", type=int, default=8, help="images in the bounded CPU sample")
    p.add_argument("--no_cpu_baseline", action="store_true")
    return p.parse_args()


MODEL_KW = dict(input_size=(32, 32, 3), aug=2, nblocks=(2, 2, 2), fx_dim=64, fw_dims=(2, 64, 2), nsteps=20,
                fx_actfn="softplus", fw_actfn="softplus")


def conv_flops_per_image_eval():
    """2*MAC of the conv stacks of the 6 blocks for one image and one drift evaluation (dense taps,
    time channel included: SURVEY.md 8.4 -> 175.1 MFLOP)."""
    from bayesde_b200.arch import FxSpec
    total = 0
    shape = (1, 32, 32, 5)
    for gi, nb in enumerate(MODEL_KW["nblocks"]):
        fx = FxSpec(0, shape, MODEL_KW["fx_dim"], "softplus")
        per = 0
        for L in fx.layers:
            pix = L["hout"] * L["wout"] if L["sd"] == 1 else L["hin"] * L["win"]  # transposed: per input pixel
            per += 2 * pix * L["k"] * L["k"] * (L["cin"] + 1) * L["cout"]
        total += nb * per
        shape = (1, shape[1] // 2, shape[2] // 2, shape[3] * 4)
    return total


def conv_bytes_per_image_eval():
    """Algorithmic HBM bytes of the conv stacks for one image and one drift evaluation with fp32 NHWC activations and no
    cross-layer fusion: every layer reads its input once and writes its output once.  Returns (fwd, bwd): bwd = data
    gradients (dy read, act'(.) operand read, dz written) + weight gradients (layer input and dy read)."""
    from bayesde_b200.arch import FxSpec
    fwd = bwd = 0
    shape = (1, 32, 32, 5)
    for gi, nb in enumerate(MODEL_KW["nblocks"]):
        fx = FxSpec(0, shape, MODEL_KW["fx_dim"], "softplus")
        f = b = 0
        n = len(fx.layers)
        for li, L in enumerate(fx.layers):
            vin, vout = L["hin"] * L["win"] * L["cin"], L["hout"] * L["wout"] * L["cout"]
            f += vin + vout + (vout if li == n - 1 else 0)      # Euler epilogue reads x_k
            b += vout + vin + vin                               # dgrad: dy, aux (act output / gx), dz
            b += vin + vout                                     # wgrad: layer input, dy
        fwd += nb * f * 4
        bwd += nb * b * 4
        shape = (1, shape[1] // 2, shape[2] // 2, shape[3] * 4)
    return fwd, bwd


def measure_tf32_peak(dev):
    """Dense TF32 tensor-core throughput measured live the way MEASURED_PEAKS.json's bf16 figure was (cuBLAS 8192^3)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 10 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_images_per_s(args, steps, warmup):
    """The oracle's torch-CPU fp32 restatement of the JAX path (the reference itself needs JAX 0.1.77,
    which cannot be installed here), all host threads, bounded sample: `cpu_batch` images x 1 MC
    sample per step, scaled to the metric's unit (images/s at nsamples samples per image)."""
    import torch
    from oracle import jax_path as jp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_batch
    net = jp.SDEBNNClassifier(batch=B, diff_coef=args.diff_coef, time_grid_mode=args.time_grid,
                              **{k: v for k, v in MODEL_KW.items()})
    params, head = net.init(0, torch.float32)
    leaves = []
    for (w0, fw) in params:
        w0.requires_grad_(True)
        leaves.append(w0)
        for lay in fw:
            for t in lay:
                t.requires_grad_(True)
                leaves.append(t)
    for pair in head:
        for t in pair:
            t.requires_grad_(True)
            leaves.append(t)
    opt = torch.optim.Adam(leaves, lr=1e-3)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 32, 32, 3, generator=g)
    onehot = torch.nn.functional.one_hot(torch.randint(0, 10, (B,), generator=g), 10).float()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        dWs = [jp.sample_increments(b.P, 20, g, torch.float32, args.time_grid) for b in net.blocks()]
        eps = torch.randn(net.feat * 10 + 10, generator=g)
        opt.zero_grad()
        loss, _, _, _ = net.loss(params, head, x, onehot, dWs, eps, 1e-3 / 45000)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t_step = sum(times) / len(times)
    value = B / (t_step * args.nsamples)
    sample = (f"{steps} timed + {warmup} warm-up train steps (fwd+bwd+Adam) of the oracle port at "
              f"{B} images x 1 MC sample, {t_step:.2f} s/step; value = {B}/(s_per_step*{args.nsamples} samples)")
    return value, cores, sample, t_step


def cpu_torch_reference_images_per_s(batch):
    """The torch-CPU leg of the north star: one training step of config 5 (torchbnn SDENet, 3x32x32, blocks 2-2-2, hidden 32, w-net
    1-128-1, midpoint dt 0.05) through the oracle's restatement of the torch path (SDENet as shipped raises: quirk Q7; torchsde is
    absent), all host threads, bounded sample: 1 warm-up + 1 timed step at `batch` images."""
    import torch
    import torch.nn.functional as F
    from oracle import torch_path as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = tp.SDENetOracle((3, 32, 32), (2, 2, 2), (1, 128, 1), sigma=0.1, hidden_width=32)
    gg = torch.Generator().manual_seed(0)
    w0 = tp.init_y_params(net.ops, net.P, gg).requires_grad_(True)
    wl = [(W.requires_grad_(True), b.requires_grad_(True)) for W, b in tp.init_w_net(net.P, net.hidden_sizes, gg, last_std=1e-2)]
    feat = int(net.output_size[0] * net.output_size[1] * net.output_size[2])
    proj = [((torch.randn(1024, feat, generator=gg) * 0.01).requires_grad_(True), torch.zeros(1024, requires_grad=True)),
            ((torch.randn(10, 1024, generator=gg) * 0.01).requires_grad_(True), torch.zeros(10, requires_grad=True))]
    leaves = [w0] + [t for pr in wl for t in pr] + [t for pr in proj for t in pr]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    xc = torch.randn(batch, 3, 32, 32, generator=gg)
    yc = torch.randint(0, 10, (batch,), generator=gg)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        I = torch.randn(20, net.P, generator=gg) * 0.05 ** 0.5
        opt.zero_grad()
        lo, lq = net.forward(xc, w0, wl, proj, I, dt=0.05)
        (F.cross_entropy(lo, yc) + 1e-4 * lq).backward()
        opt.step()
        ts.append(time.perf_counter() - t0)
    sample = (f"1 warm-up + 1 timed train step (fwd+bwd+Adam, 40 drift evaluations) of oracle/torch_path.py at {batch} images, "
              f"{ts[-1]:.2f} s/step")
    return batch / ts[-1], cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warmup = 1 if args.warmup > 0 else 0
    v, cores, sample, t_step = cpu_reference_images_per_s(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "config4: SDEBNN 32x32x3 CIFAR-10-shape synthetic, aug 2, nblocks 2-2-2, fx_dim 64, "
                        "fw_dims 2-64-2, softplus, nsteps 20 (19 drift evaluations/block), euler_maruyama",
            "batch_per_gpu": args.batch, "nsamples": args.nsamples, "diff_coef": args.diff_coef,
            "time_grid": args.time_grid, "conv_math": args.math, "remat": bool(getattr(args, "remat", False)), "parallelism": f"dp{args.gpus} (batch sharded)",
            "l2_policy": "per-step working set (checkpoints+activations, >2 GB) exceeds the 126 MB L2; no flush"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bayesde_b200 import _fused, _lib
    from bayesde_b200.classification import SDEBNNClassifier, Trainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep stdout to the one JSON line: NCCL writes its version banner to fd 1 from C, so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    model = SDEBNNClassifier(nsamples=args.nsamples, diff_coef=args.diff_coef, time_grid=args.time_grid,
                             math=args.math, remat=args.remat, w_init=1e-2, b_init=1e-2, seed=0, device=dev, **MODEL_KW)
    trainer = Trainer(model, lr=1e-3, kl_coef=1e-3, train_size=45000)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_x = [torch.randn(B, 32, 32, 3, generator=g).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 10, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    seed_base = [0]

    def step_resident(i):
        trainer.step(dev_x[i % n_host], dev_y[i % n_host], seed=seed_base[0] + i)

    last = {}

    def step_e2e(i):
        x = host_x[i % n_host].to(dev, non_blocking=True)
        y = host_y[i % n_host].to(dev, non_blocking=True)
        loss, _, _ = trainer.step(x, y, seed=seed_base[0] + i)
        last["loss"] = loss.item()  # device -> host read of the step's result

    for i in range(args.warmup):
        step_resident(i)
    seed_base[0] = 1000

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _fused.PROFILE = {}
    n0 = L.bsde_launch_count()
    ms = timed(step_resident, args.steps)
    launches = L.bsde_launch_count() - n0
    prof = _fused.PROFILE
    _fused.PROFILE = None
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ktime = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in prof.items()}

    seed_base[0] = 2000
    ms_e2e = timed(step_e2e, args.steps)

    images = B * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    fl_eval = conv_flops_per_image_eval()
    tf32_peak = measure_tf32_peak(dev) if args.math == "tc" else None
    by_fwd, by_bwd = conv_bytes_per_image_eval()
    nit = MODEL_KW["nsteps"] - 1
    alg_flops_step = 3.0 * fl_eval * nit * B * args.nsamples  # fwd + dgrad + wgrad (recompute not counted)
    conv_ms = (ktime.get("k2_fwd", 0.0) + ktime.get("k3_bwd", 0.0)) / args.steps
    hbm_bytes_step = (by_fwd + by_bwd) * nit * B * args.nsamples
    hbm_achieved = hbm_bytes_step / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else 0.0
    achieved = alg_flops_step / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # K1 (weight-SDE) against the 48*P HBM model of SURVEY.md 8.4 (fwd) and 2x for the reverse pass
    sumP = sum(blk[0].numel() for blk in model.block_params())
    k1_bytes = 3 * 48.0 * sumP * nit * args.nsamples
    k1_ms = (ktime.get("k1_fwd", 0.0) + ktime.get("k1_bwd", 0.0)) / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.math == "fp32" else "tf32 (tcgen05 kind::tf32 operands, f32 accumulate)", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (32 * 32 * 3 * 4 + 8),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # The conv stack sits on the MEMORY side of the ridge: 10.2 PFLOP / 230 GB per step = 44 FLOP/B against a ridge of
        # (TF32 peak 748 TFLOP/s) / (6.55 TB/s) = 114 FLOP/B (byte floor 35 ms > flop floor 13.7 ms), so the bounding roofline of the
        # dominant kernels (conv_win / wgrad_win, 79 % of the step in the ncu launch list) is HBM; the tensor-pipe view is kept beside it.
        "roofline": {"bound": "hbm", "kernel": "K2/K3 time-dependent conv stack (conv_win_kernel / wgrad_win_kernel + conv_tc / wgrad_tc "
                                                "fallbacks), all launches of a step",
                     "achieved": hbm_achieved, "peak": peak_hbm, "unit": "GB/s", "frac": hbm_achieved / peak_hbm,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel with the largest share of the
                     # step (wgrad_win_kernel, 29 %: 5->64 weight gradient, 1024 images 32x32), ncu --set full capture
                     # profiles/r01j_wgrad_win_full_summary.csv: 289.8 MB read + 3.9 MB written = the algorithmic bytes
                     # (dY 268.4 MB + input 21.0 MB read once, partial tiles 3.5 MB); the eight group-0 conv_win launches
                     # likewise move exactly their algorithmic bytes (profiles/r01k_conv_win_full_summary.csv)
                     "traffic": 293.7e6,
                     "algorithmic_bytes_per_step": hbm_bytes_step, "kernel_ms_per_step": conv_ms,
                     "share_of_step": conv_ms / (ms / args.steps),
                     "byte_floor_ms": hbm_bytes_step / (peak_hbm * 1e9) * 1e3,
                     "flop_floor_ms": alg_flops_step / ((tf32_peak or peak_tf) * 1e12) * 1e3,
                     "arithmetic_intensity_flop_per_byte": alg_flops_step / hbm_bytes_step,
                     "ridge_flop_per_byte": (tf32_peak or peak_tf) * 1e12 / (peak_hbm * 1e9),
                     "note": "algorithmic bytes: every layer reads its fp32 NHWC input once and writes its output once; dgrad also "
                             "reads the activation it differentiates; wgrad reads layer input + dy.  achieved = those bytes / "
                             "CUDA-event time of the x-path calls (K2 forward + K3 reverse) on the launching stream"},
        "roofline_tensor": {"bound": "tensor", "kernel": "the same K2/K3 conv launches (tcgen05 kind::tf32)",
                            "achieved": achieved, "peak": tf32_peak or peak_tf, "unit": "TFLOP/s", "frac": achieved / (tf32_peak or peak_tf),
                            "peak_source": ("measured live in this run: cuBLAS 8192^3 TF32 matmul (the kernels issue tcgen05 kind::tf32; "
                                            "MEASURED_PEAKS.json holds no TF32 figure, SURVEY 8.4)" if tf32_peak else peak_src),
                            "peak_bf16_sustained": peak_tf, "frac_of_bf16_peak": achieved / peak_tf,
                            "algorithmic_flops_per_step": alg_flops_step},
        "roofline_k1": {"bound": "hbm", "kernel": "K1 wsde_fwd/bwd (cooperative persistent)",
                        "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0, "peak": peak_hbm,
                        "unit": "GB/s", "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9 / peak_hbm) if k1_ms > 0 else 0.0,
                        "algorithmic_bytes_per_step": k1_bytes, "kernel_ms_per_step": k1_ms,
                        # ncu --set full of wsde_fwd_kernel (profiles/r01m_wsde_full_summary.csv, group-0 block, S = 8, 19 iterations):
                        # 3.0 MB read + 3.0-3.8 MB written per launch against 594 MB algorithmic -- the state and the f_w parameters
                        # stay in the 126 MB L2, so this kernel is NOT HBM-bound: issue active 18 %, 8 warps per SM, one grid barrier
                        # per iteration (29 us per iteration, of which ~1.4 us per sample is arithmetic)
                        "traffic": 6.0e6,
                        "note": "48*P B per (sample, block, iteration) fwd, 2x bwd (SURVEY 8.4 model); measured DRAM traffic is 1 % of it "
                                "(L2-resident), the kernel is latency / barrier bound, so this fraction is a model figure, not a bound"},
        "kernel_ms_per_step": {k: v / args.steps for k, v in ktime.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_images_per_s(args, 2, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        try:  # the second CPU leg the north star names (torch front end, config 5); reported beside, not a ratio input
            v5, cores5, sample5 = cpu_torch_reference_images_per_s(args.cpu_batch)
            line["cpu_baseline_torch"] = {"value": v5, "unit": "images/s (config 5: torchbnn SDENet, midpoint dt 0.05, 1 weight sample)",
                                          "cores": cores5, "kind": "port", "sample": sample5}
        except Exception as e:  # noqa: BLE001 -- a reporting extra must not cost the bench line
            line["cpu_baseline_torch"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

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
    p.add_argument("--math", default=os.environ.get("BSDE_MATH", "tc"), choices=["fp32", "tc", "bf16"])
    p.add_argument("--shard", default="batch", choices=["batch", "samples"],
                   help="batch: B images x S samples per rank (weak scaling, default); samples: the S x B work of one step on a "
                        "samples x batch grid of ranks (strong scaling, SURVEY 8.5)")
    p.add_argument("--no_overlap", action="store_true", help="serial gradient all-reduce after backward (default: per-block buckets "
                                                             "launched on a side stream during the reverse pass)")
    p.add_argument("--remat", action="store_true", help="recompute conv intermediates in the reverse pass")
    p.add_argument("--time_grid", default="reference", choices=["reference", "uniform"])
    p.add_argument("--diff_coef", type=float, default=0.2)
    p.add_argument("--cpu_batch", type=int, default=32, help="images in the bounded CPU sample (SURVEY 8.4: B = 32)")
    p.add_argument("--no_fp32_leg", action="store_true", help="skip the fp32-class (CUDA-core conv) throughput leg")
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--no_config5", action="store_true", help="skip the config-5 (torchbnn SDENet) leg that the default run appends as `config5`")
    p.add_argument("--config", type=int, default=4, choices=[4, 5],
                   help="4: the north-star JAX-front-end SDEBNN classifier (default); 5: BASELINE configs[4], the torchbnn SDENet "
                        "(3x32x32, blocks 2-2-2, hidden 32, w-net 1-128-1, midpoint dt 0.05, inhomogeneous) with the batch sharded over ranks")
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
    """The oracle's torch-CPU fp32 restatement of the JAX path (the reference itself needs JAX 0.1.77, which cannot be
    installed here), all host threads, bounded sample per SURVEY 8.4: `cpu_batch` (32) images x 1 MC sample per train step
    (forward + backward + Adam), `warmup` untimed + `steps` timed steps.  The metric counts images/s with every image integrated
    under `nsamples` weight paths, and the samples are independent repeats of the same work (script :218 vmaps over rngs), so
    value = cpu_batch / (s_per_step * nsamples): extrapolated in S, stated in `ran`."""
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
    ran = {"batch": B, "nsamples": 1, "steps": steps, "warmup": warmup, "extrapolated": True,
           "extrapolation": f"x 1/{args.nsamples} in the MC-sample count (independent repeats of the measured work)"}
    return value, cores, sample, t_step, ran


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
    steps = max(5, min(args.steps, 8))   # >= 5 timed steps; bounded so the arm ends within a few minutes
    warmup = 3 if args.warmup >= 3 else max(1, args.warmup)
    v, cores, sample, t_step, ran = cpu_reference_images_per_s(args, steps, warmup)
    cfg = workload_config(args)
    # what this arm actually ran (the GPU arm's batch_per_gpu / nsamples are the metric's definition, not this arm's sample)
    cfg.update({"conv_math": "fp32 (torch CPU)", "parallelism": f"host threads x{cores}", "reference_arm_ran": ran,
                "l2_policy": "n/a (CPU)"})
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    shard = getattr(args, "shard", "batch")
    par = (f"dp{args.gpus} (batch sharded: {args.batch} images x {args.nsamples} samples per rank)" if shard == "batch" else
           f"samples x batch grid over {args.gpus} ranks (SURVEY 8.5): {args.nsamples} samples and {args.batch} images in total")
    return {"workload": "config4: SDEBNN 32x32x3 CIFAR-10-shape synthetic, aug 2, nblocks 2-2-2, fx_dim 64, "
                        "fw_dims 2-64-2, softplus, nsteps 20 (19 drift evaluations/block), euler_maruyama",
            "batch_per_gpu": args.batch, "nsamples": args.nsamples, "diff_coef": args.diff_coef,
            "time_grid": args.time_grid, "conv_math": args.math, "remat": bool(getattr(args, "remat", False)), "parallelism": par,
            "shard": shard,
            "l2_policy": "per-step working set (checkpoints+activations, >2 GB) exceeds the 126 MB L2; no flush"}


def fused_floor_bytes_per_image_eval():
    """HBM byte floor of a block evaluation if conv1->conv2 and convT->conv4 were fused pairs: per block
    the state x is read and written once (fp32) and only the half-resolution 64-channel tensor crosses HBM (written, read);
    reverse pass with recompute: x read, gx read + written, half-resolution activation and its gradient written + read."""
    from bayesde_b200.arch import FxSpec
    fwd = bwd = 0
    shape = (1, 32, 32, 5)
    for gi, nb in enumerate(MODEL_KW["nblocks"]):
        fx = FxSpec(0, shape, MODEL_KW["fx_dim"], "softplus")
        vx = fx.H * fx.W * fx.C
        vmid = fx.layers[1]["hout"] * fx.layers[1]["wout"] * fx.layers[1]["cout"]
        fwd += nb * 4 * (2 * vx + 2 * vmid)
        bwd += nb * 4 * (3 * vx + 4 * vmid)
        shape = (1, shape[1] // 2, shape[2] // 2, shape[3] * 4)
    return fwd, bwd


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bayesde_b200 import _fused, _lib
    from bayesde_b200.classification import SDEBNNClassifier, Trainer, sample_shard

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

    # ---- decomposition.  batch (default, weak scaling): every rank integrates all S weight paths on its own B images.
    # samples (strong scaling): the S x B work of ONE step is laid on a samples x batch grid -- S/G samples per rank
    # on the whole batch while G <= S, beyond that ranks of a sample group split the batch; every sample's KL is counted once.
    B = args.batch
    if args.shard == "samples":
        s_off, s_loc, b_grp, grp = sample_shard(args.nsamples, world, rank)
        if B % grp:
            raise SystemExit(f"--batch {B} is not divisible by the {grp} ranks of a sample group")
        B_loc = B // grp
    else:
        s_off, s_loc, b_grp, grp, B_loc = 0, args.nsamples, rank, 1, B
    mk = dict(nsamples=s_loc, sample_offset=s_off, total_samples=args.nsamples, diff_coef=args.diff_coef,
              time_grid=args.time_grid, remat=args.remat, w_init=1e-2, b_init=1e-2, seed=0, device=dev, **MODEL_KW)
    model = SDEBNNClassifier(math=args.math, **mk)
    trainer = Trainer(model, lr=1e-3, kl_coef=1e-3, train_size=45000, overlap=not args.no_overlap)
    # batch mode: different images per rank; samples mode: the SAME global batch on every rank, a group's ranks take slices of it
    g = torch.Generator().manual_seed(1234 + (rank if args.shard == "batch" else 0))
    n_host = 4
    lo = b_grp * B_loc if args.shard == "samples" else 0
    host_x = [torch.randn(B, 32, 32, 3, generator=g)[lo:lo + B_loc].contiguous().pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 10, (B,), generator=g)[lo:lo + B_loc].contiguous().pin_memory() for _ in range(n_host)]
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
    last = {}

    def make_steps(tr):
        def step_resident(i):
            out = tr.step(dev_x[i % n_host], dev_y[i % n_host], seed=seed_base[0] + i)
            last["dev"] = out

        def step_e2e(i):
            x = host_x[i % n_host].to(dev, non_blocking=True)
            y = host_y[i % n_host].to(dev, non_blocking=True)
            loss, _, _ = tr.step(x, y, seed=seed_base[0] + i)
            last["loss"] = loss.item()  # device -> host read of the step's result
        return step_resident, step_e2e

    step_resident, step_e2e = make_steps(trainer)
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
    loss_last = [float(v) for v in last["dev"]]   # (loss, nll, kl) of the last timed step (read after the timed region)

    seed_base[0] = 2000
    ms_e2e = timed(step_e2e, args.steps)

    imgs_step = B * world if args.shard == "batch" else B
    images = imgs_step * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)

    # ---- fp32-class leg (CUDA-core fp32 convolutions, the arithmetic of the reference) on the same workload, and the same-seed
    # loss of the two arithmetics on one un-timed step from identical parameters: the benchmarked step is a sane computation
    fp32_leg = None
    if args.math != "fp32" and not args.no_fp32_leg:
        seed_cmp = 4242
        with torch.no_grad():
            loss_fast = float(model.loss(dev_x[0], dev_y[0], trainer.kl_scale, rng=seed_cmp)[0])
        m32 = SDEBNNClassifier(math="fp32", **mk)
        m32.load_state_dict(model.state_dict())
        with torch.no_grad():
            loss_32 = float(m32.loss(dev_x[0], dev_y[0], trainer.kl_scale, rng=seed_cmp)[0])
        t32 = Trainer(m32, lr=1e-3, kl_coef=1e-3, train_size=45000, overlap=not args.no_overlap)
        st32, _ = make_steps(t32)
        st32(0)
        n32 = max(1, min(2, args.steps))
        ms32 = timed(st32, n32)
        fp32_leg = {"value": imgs_step * n32 / (ms32 * 1e-3), "unit": UNIT, "steps": n32, "warmup": 1, "ms_per_step": ms32 / n32,
                    "conv_math": "fp32 (CUDA-core FMA, csrc/tconv_simt.cu)",
                    "same_seed_loss": {"fast": loss_fast, "fp32": loss_32,
                                       "rel_diff": abs(loss_fast - loss_32) / max(abs(loss_32), 1e-30)}}
        del m32, t32

    # ---- BASELINE configs[4] beside it (torch front end, same ranks, batch sharded through DistributedDataParallel): a second,
    # driver-visible throughput at every N.  Not part of `value`.
    c5 = None
    if not args.no_config5 and args.shard == "batch":
        try:
            c5 = measure_config5(args, dev, rank, world, local)
        except Exception as e:  # noqa: BLE001 -- an extra leg must not cost the bench line
            c5 = {"unavailable": repr(e)}

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
    peak_bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    fl_eval = conv_flops_per_image_eval()
    tf32_peak = measure_tf32_peak(dev) if args.math == "tc" else None
    by_fwd, by_bwd = conv_bytes_per_image_eval()
    ff_fwd, ff_bwd = fused_floor_bytes_per_image_eval()
    nit = MODEL_KW["nsteps"] - 1
    units = B_loc * s_loc                                      # image-samples one rank integrates per step
    alg_flops_step = 3.0 * fl_eval * nit * units               # fwd + dgrad + wgrad (recompute not counted): 9.98 GFLOP/image-sample
    conv_ms = (ktime.get("k2_fwd", 0.0) + ktime.get("k3_bwd", 0.0)) / args.steps
    hbm_bytes_step = (by_fwd + by_bwd) * nit * units
    floor_bytes_step = (ff_fwd + ff_bwd) * nit * units
    hbm_achieved = hbm_bytes_step / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else 0.0
    achieved = alg_flops_step / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    if args.math == "tc":
        t_peak, t_src = tf32_peak, ("measured live in this run: cuBLAS 8192^3 TF32 matmul (the kernels issue tcgen05 kind::tf32; "
                                    "MEASURED_PEAKS.json holds no TF32 figure, SURVEY 8.4)")
    else:
        t_peak, t_src = peak_bf16, peak_src
    # ncu DRAM traffic of the dominant kernel: read from the committed summary of the capture, never typed in
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            tjd = json.load(open(tj)).get(args.math)
            if tjd:
                traffic, traffic_src = tjd.get("dram_bytes_per_launch"), tjd.get("source")
        except Exception:  # noqa: BLE001
            pass
    # K1 (weight-SDE) against the 48*P HBM model of SURVEY.md 8.4 (fwd) and 2x for the reverse pass
    sumP = sum(blk[0].numel() for blk in model.block_params())
    k1_bytes = 3 * 48.0 * sumP * nit * s_loc
    k1_ms = (ktime.get("k1_fwd", 0.0) + ktime.get("k1_bwd", 0.0)) / args.steps
    dtype = {"fp32": "f32", "tc": "tf32 (tcgen05 kind::tf32 operands, f32 accumulate)",
             "bf16": "bf16 (tcgen05 kind::f16 operands and bf16 storage of the 64-channel intermediates; x, dx, accumulators, "
                     "all parameter gradients f32)"}[args.math]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.shard == "batch" else "strong",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B_loc * (32 * 32 * 3 * 4 + 8),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "loss": {"last_timed_step": loss_last[0], "nll": loss_last[1], "kl": loss_last[2], "last_e2e_step": last.get("loss"),
                 "finite": bool(all(map(lambda v: v == v and abs(v) != float("inf"), loss_last + [last.get("loss", 0.0)])))},
        "value_fp32": fp32_leg["value"] if fp32_leg else (value if args.math == "fp32" else None),
        "fp32_leg": fp32_leg,
        # the per-step conv is a dense contraction => tensor-core bound; achieved = algorithmic FLOPs (9.98 GFLOP per
        # image-sample and training step = 3 x 175.1 MFLOP x 19 iterations) / CUDA-event time of the K2 + K3 calls on the launching stream
        "roofline": {"bound": "tensor", "kernel": "K2/K3 time-dependent conv stack (conv_win_kernel / wgrad_win_kernel + fallbacks), all "
                                                  "launches of a step",
                     "achieved": achieved, "peak": t_peak, "unit": "TFLOP/s", "frac": achieved / t_peak,
                     "peak_source": t_src, "peak_bf16_sustained": peak_bf16, "frac_of_bf16_peak": achieved / peak_bf16,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flops_per_step": alg_flops_step, "flops_per_image_sample": 3.0 * fl_eval * nit,
                     "image_samples_per_step": units, "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / (ms / args.steps),
                     "flop_floor_ms": alg_flops_step / (t_peak * 1e12) * 1e3},
        # the explanation beside it: with fp32 NHWC activations and one launch per layer the stack sits on the memory side of the ridge
        "roofline_hbm": {"bound": "hbm", "kernel": "the same K2/K3 launches", "achieved": hbm_achieved, "peak": peak_hbm, "unit": "GB/s",
                         "frac": hbm_achieved / peak_hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "per_layer_bytes_per_step": hbm_bytes_step, "per_layer_byte_floor_ms": hbm_bytes_step / (peak_hbm * 1e9) * 1e3,
                         "fused_block_floor_bytes_per_step": floor_bytes_step,
                         "fused_block_floor_ms": floor_bytes_step / (peak_hbm * 1e9) * 1e3,
                         "note": "per_layer: every layer reads its input once and writes its output once at the storage width of this "
                                 "run's arithmetic (fp32 here unless conv_math = bf16); dgrad also reads the activation it differentiates; "
                                 "wgrad reads layer input + dy.  fused_block_floor: conv1->conv2 and convT->conv4 as fused pairs -- only x, "
                                 "gx and the half-resolution 64-channel tensor cross HBM (VERDICT r1 item 5)"},
        "roofline_k1": {"bound": "hbm", "kernel": "K1 wsde_fwd/bwd (cooperative persistent)",
                        "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0, "peak": peak_hbm,
                        "unit": "GB/s", "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9 / peak_hbm) if k1_ms > 0 else 0.0,
                        "algorithmic_bytes_per_step": k1_bytes, "kernel_ms_per_step": k1_ms, "traffic": None,
                        "note": "48*P B per (sample, block, iteration) fwd, 2x bwd (SURVEY 8.4 model); the state and the f_w parameters stay "
                                "in L2 (ncu: profiles/r01m_wsde_full_summary.csv, ~1 % of the model bytes reach DRAM), so this fraction is a "
                                "model figure, not a bound"},
        "kernel_ms_per_step": {k: v / args.steps for k, v in ktime.items()},
    }
    if c5 is not None:
        line["config5"] = c5
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _, ran = cpu_reference_images_per_s(args, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ran": ran}
        try:  # the second CPU leg the north star names (torch front end, config 5); reported beside, not a ratio input
            v5, cores5, sample5 = cpu_torch_reference_images_per_s(8)
            line["cpu_baseline_torch"] = {"value": v5, "unit": "images/s (config 5: torchbnn SDENet, midpoint dt 0.05, 1 weight sample)",
                                          "cores": cores5, "kind": "port", "sample": sample5}
        except Exception as e:  # noqa: BLE001 -- a reporting extra must not cost the bench line
            line["cpu_baseline_torch"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config5(args):
    """`bench.py --config 5`: the config-5 line alone (see measure_config5)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = measure_config5(args, dev, rank, world, local, full=True)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_config5(args, dev, rank, world, local, full=False):
    """BASELINE configs[4]: torchbnn SDENet training step (examples/torch/sdebnn_classification.py:34-46: cross-entropy +
    kl_coeff * logqp, Adam) at `--batch` images per GPU, midpoint dt 0.05 (20 steps, 40 evaluations of the 26-conv y_net),
    batch sharded over ranks: every rank integrates the SAME weight path (same Philox key, counter-based) on its own images;
    the one collective is the gradient all-reduce, here torch's DistributedDataParallel (bucketed, overlapped with backward).
    Called on every rank (the process group, if any, is already up); returns the JSON object on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from bayesde_b200 import _lib, torch_models as tm

    L = _lib.lib()
    torch.manual_seed(0)
    hidden = 32
    model = tm.SDENet(input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 128, 1), hidden_width=hidden, sigma=0.1,
                      inhomogeneous=True, math="tc" if args.math != "fp32" else "fp32", remat=args.remat or args.batch > 256).to(dev)
    with torch.no_grad():   # N(0, 1e-2) on the zero-initialised last w-net layer, else the w drift is trivially -w
        last = model.w_net.linears()[-1]
        last.weight.normal_(0, 1e-2), last.bias.normal_(0, 1e-2)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_x = [torch.randn(B, 3, 32, 32, generator=g).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 10, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    last_loss = {}

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        logits, logqp = net(x, dt=0.05, method="midpoint")
        loss = F.cross_entropy(logits, y) + 1e-4 * logqp
        loss.backward()
        opt.step()
        return loss.detach()

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

    def step_resident(i):
        last_loss["dev"] = step(dev_x[i % n_host], dev_y[i % n_host])

    def step_e2e(i):
        x = host_x[i % n_host].to(dev, non_blocking=True)
        y = host_y[i % n_host].to(dev, non_blocking=True)
        last_loss["e2e"] = step(x, y).item()

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.bsde_launch_count()
    ms = timed(step_resident, args.steps)
    launches = L.bsde_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    del net, opt
    if rank != 0:
        return None
    tf32_peak = measure_tf32_peak(dev)
    flops_eval = sum(2 * (l["hout"] * l["wout"] if not l["transpose"] else l["hin"] * l["win"]) * 9 * (l["cin"] + 1) * l["cout"]
                     for l in model.y_net.layers)
    alg = 3.0 * flops_eval * 40 * B     # fwd + dgrad + wgrad of 40 drift evaluations, per rank
    images = B * world * args.steps
    achieved = alg / (ms / args.steps * 1e-3) / 1e12
    loss_v = float(last_loss["dev"])
    line = {"metric": "SDENet (torchbnn) train images/s (3x32x32, midpoint dt 0.05)", "value": images / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (tcgen05 kind::tf32 operands, f32 accumulate)", "data": "synthetic",
            "config": {"workload": f"config5: torchbnn SDENet 3x32x32, blocks 2-2-2, hidden {hidden}, w-net 1-128-1, sigma 0.1, midpoint "
                                   f"dt 0.05 (40 drift evaluations), inhomogeneous, P = {model.params_size}",
                       "batch_per_gpu": B, "remat": bool(model.remat), "parallelism": f"dp{world} (batch sharded, DistributedDataParallel)",
                       "l2_policy": "per-step working set exceeds the 126 MB L2; no flush"},
            "e2e": {"value": images / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * (3 * 32 * 32 * 4 + 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "loss": {"last_timed_step": loss_v, "last_e2e_step": last_loss.get("e2e"), "finite": loss_v == loss_v and abs(loss_v) != float("inf")},
            "roofline": {"bound": "tensor", "kernel": "the y_net conv chain (bsde_chain_fwd / bsde_chain_bwd launches) -- against the WHOLE step time",
                         "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                         "peak_source": "measured live in this run: cuBLAS 8192^3 TF32 matmul", "traffic": None,
                         "algorithmic_flops_per_step": alg, "flops_per_image_eval": flops_eval}}
    if full and world == 1 and not args.no_cpu_baseline:
        try:
            v5, cores5, sample5 = cpu_torch_reference_images_per_s(8)
            line["cpu_baseline"] = {"value": v5, "unit": UNIT, "cores": cores5, "kind": "port", "sample": sample5}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"unavailable": repr(e)}
    return line


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == 5:
        run_config5(a)
    else:
        run_ours(a)

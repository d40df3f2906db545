#!/usr/bin/env python
"""bench.py -- SVI samples/sec on BASELINE.json configs[1]:
iVAE 2D rot+trans invariant, synthetic 28x28 Bernoulli images, latent_dim=2,
FC encoder / spatial FC decoder, batch 512 per GPU (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full SVItrainer mini-batch step (encoder fwd, latent sample,
affine fold, spatial decoder fwd+bwd, ELBO, encoder bwd, [all-reduce], Adam)
on one batch of synthetic input.  Prints ONE JSON line (rank 0).

  value : whole-job samples/s with inputs already resident in HBM
  e2e   : same metric through the public API (`SVItrainer.train(loader)` over
          pinned host batches): host -> device copy of every batch and
          device -> host read of every step's loss inside the timed region
  roofline     : dominant kernel, timed alone with CUDA events (live)
  cpu_baseline : the oracle port (oracle/svi_port.py, torch CPU fp32, all host
                 threads) timed on a bounded sample of the same workload
  --impl reference : times that CPU port as the reference arm
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

H = W = 28
LATENT = 2
INV = ['r', 't']
BATCH = 512
POOL = 96  # input batches kept in HBM: 96 x 1.6 MB = 154 MB > 126 MB L2

# algorithmic work (SURVEY.md 8d / DESIGN.md)
FLOP_PER_ROW_FWD = 66304           # 2*(2*128 + 2*128*128 + 128) per pixel-row
FLOP_PER_ROW_STEP = 3 * FLOP_PER_ROW_FWD


def synth_batches(n_batches, batch, seed=0):
    """SURVEY 8(d) cfg2 data: rotated / shifted anisotropic Gaussian blobs,
    Bernoulli-sampled."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    n = n_batches * batch
    th = (torch.rand(n, generator=g) * 2 - 1) * math.pi / 3
    t = (torch.rand(n, 2, generator=g) * 2 - 1) * 0.1
    xx = torch.linspace(-1, 1, H)
    yy = torch.linspace(1, -1, W)
    gx, gy = torch.meshgrid(xx, yy, indexing="ij")
    gx = gx[None] - t[:, 0, None, None]
    gy = gy[None] - t[:, 1, None, None]
    c, s = torch.cos(th)[:, None, None], torch.sin(th)[:, None, None]
    u = c * gx + s * gy
    v = -s * gx + c * gy
    p = torch.exp(-(u ** 2 / (2 * 0.15 ** 2) + v ** 2 / (2 * 0.45 ** 2)))
    x = (torch.rand(n, H, W, generator=g) < p).float()
    return x.reshape(n_batches, batch, H, W)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "tf": p.get("bf16_tflops", 1590.0),
                "tf_sustained": p.get("bf16_tflops_sustained", 1400.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.stop = False
        self.th = None

    def _loop(self):
        while not self.stop:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True, text=True,
                    timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(batch, max_seconds, steps=None, warmup=0):
    """Time the oracle port's SVI step (fwd + bwd + Adam) on host cores."""
    import torch
    from oracle import svi_port as sp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = sp.Cfg((H, W), LATENT, INV)
    port = sp.SVIPort(cfg, seed=1)
    x = synth_batches(4, batch, seed=0)
    for i in range(warmup):
        port.step(x[i % 4])
    t0 = time.perf_counter()
    n = 0
    loss = None
    while True:
        loss = port.step(x[n % 4])
        n += 1
        el = time.perf_counter() - t0
        if steps is not None and n >= steps:
            break
        if steps is None and (el >= max_seconds or n >= 16):
            break
    el = time.perf_counter() - t0
    return {"value": batch * n / el, "steps": n, "seconds": el, "cores": cores,
            "loss_per_sample": loss / batch, "ms_per_step": 1e3 * el / n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # Each step is one batch of the workload; if K + W batches of 512 would take more than ~150 s on this
    # box's host cores, a step becomes a smaller batch of the same data (CPU samples/s is flat in the
    # batch size) so the run stays within a few minutes.
    probe = cpu_port_throughput(BATCH, 0, steps=1, warmup=1)
    total = (args.steps + args.warmup) * probe["seconds"]
    step_batch = BATCH
    if total > 150.0:
        step_batch = max(32, int(BATCH * 150.0 / total) // 32 * 32)
    r = cpu_port_throughput(step_batch, 0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "SVI samples/sec (28x28 iVAE rot+trans)",
        "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "iVAE 2D rot+trans, 28x28 Bernoulli, latent_dim=2, fc enc / "
                               "spatial fc dec, batch=512 (BASELINE configs[1]); CPU port of "
                               "the reference path, one batch of {} per step".format(step_batch)},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"],
                         "kind": "port",
                         "sample": "{} steps of batch {} (oracle/svi_port.py, torch CPU fp32; "
                                   "real Pyro is not installable offline)".format(r["steps"], step_batch)},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_kernel_alone(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = "cuda:{}".format(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    import pyroved_b200 as pv
    from pyroved_b200 import _lib, ops

    model = pv.models.iVAE((H, W), latent_dim=LATENT, invariances=INV, seed=1, device=dev)
    trainer = pv.trainers.SVItrainer(model, seed=1, device=dev)
    svi = trainer.svi

    host = synth_batches(POOL, BATCH, seed=1000 + rank)          # [POOL,B,H,W] this rank's shard
    host_pinned = host.pin_memory()
    pool = host.to(dev)                                           # resident in HBM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also triggers graph capture) ---------------------------------
    W_ = max(args.warmup, 3)
    lc0 = _lib.lib().pvb_launch_count()
    svi.step(pool[0])                                              # eager: counts launches
    launches_per_step = _lib.lib().pvb_launch_count() - lc0
    for i in range(1, W_):
        svi.step(pool[i % POOL], _sync=False)
    barrier()

    # ---- device-resident throughput ("value") ------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            svi.step(pool[(W_ + i) % POOL], _sync=False)
        e1.record()
        barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    loss_last = float(svi.flat.loss.item()) / BATCH / world
    # ---- end to end through the public API ------------------------------------------
    # `trainer.train(loader)` (= one epoch of SVItrainer.step): host batches in pinned memory,
    # every step copies its batch H2D and its loss D2H inside the timed region.
    nb = min(args.steps, POOL)
    loader = pv.utils.TensorBatchLoader(host_pinned[:nb].reshape(nb * BATCH, H, W),
                                        batch_size=BATCH, shuffle=False, pin_memory=True)
    epochs = (args.steps + nb - 1) // nb
    trainer.train(loader)                                           # warm-up epoch (staging buffers)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(epochs):
        loss_f = trainer.train(loader) * BATCH                      # mean loss per sample -> per batch
    e3.record()
    barrier()
    e2e_steps = epochs * nb
    t_e2e = e2.elapsed_time(e3) * 1e-3 * args.steps / e2e_steps     # normalised to args.steps

    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        pk = peaks()
        prog = next(iter(svi.programs.values()))
        R = BATCH * H * W
        use_tc = getattr(prog, "use_tc", False)
        kernels = {}
        from pyroved_b200.nets.fc import linear_layers
        dec = model.decoder
        L = linear_layers(dec.fc_layers)
        # fused grid + affine + first layer kernel (HBM-write bound when it materialises h0)
        if not use_tc:
            t_h0 = time_kernel_alone(lambda: ops.sdec_h0_fwd(prog.dec.Uv, prog.dec.h0, H, W, 2))
            kernels["pvb_sdec_h0_fwd"] = {
                "bound": "hbm", "achieved": R * 128 * 4 / t_h0 / 1e9, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": R * 128 * 4 / t_h0 / 1e9 / pk["hbm_gbs"], "traffic": None,
                "us": t_h0 * 1e6}
            t_mm = time_kernel_alone(lambda: ops.linear_fwd(
                prog.dec.h0, L[0].weight.data, L[0].bias.data, "tanh", out=prog.dec.dmlp.h[0]))
            fl = 2.0 * R * 128 * 128
            kernels["sgemm_kernel(linear_fwd 128x128)"] = {
                "bound": "tensor", "achieved": fl / t_mm / 1e12, "peak": pk["tf"],
                "unit": "TFLOP/s", "frac": fl / t_mm / 1e12 / pk["tf"], "traffic": None,
                "us": t_mm * 1e6,
                "note": "fp32 SIMT generic path (no tensor cores); tcgen05 kernel not active"}
            roof = dict(kernels["sgemm_kernel(linear_fwd 128x128)"])
            roof["kernel"] = "sgemm_kernel(linear_fwd 128x128)"
        else:
            def tc_once():
                ops.sdec_tc_step(prog.dec.Uv, prog.x, None, L[0].weight.data, L[0].bias.data,
                                 L[1].weight.data, L[1].bias.data, dec.out.weight.data,
                                 dec.out.bias.data, prog.dec.rowll, prog.loc, prog.dec.gUv_part,
                                 prog.dec.wgrad_part, prog.dec.I, prog.B, H, W, 2, "bernoulli", True,
                                 0.5, True)
            t_tc = time_kernel_alone(tc_once)
            fl = float(FLOP_PER_ROW_STEP) * R
            roof = {"kernel": "pvb_sdec_tc_step (fused fwd+bwd spatial decoder)",
                    "bound": "tensor", "achieved": fl / t_tc / 1e12, "peak": pk["tf"],
                    "unit": "TFLOP/s", "frac": fl / t_tc / 1e12 / pk["tf"], "traffic": None,
                    "us": t_tc * 1e6}
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    roof["traffic"] = json.load(f)[roof["kernel"]]["dram_bytes_per_launch"]
            except Exception:
                roof["traffic"] = None
            kernels[roof["kernel"]] = roof
        roof["peak_source"] = pk["src"] + " (MEASURED_PEAKS.json burst: kernel timed alone)"
        if os.environ.get("PVB_BENCH_SKIP_CPU") == "1":   # profiling runs only
            cpu = {"value": None, "cores": 0, "steps": 0, "seconds": 0.0}
        else:
            cpu = cpu_port_throughput(BATCH, 12.0)
        total = BATCH * world * args.steps
        line = {
            "metric": "SVI samples/sec (28x28 iVAE rot+trans)",
            "value": total / t_dev, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 tensor-core operands, fp32 accumulate)" if use_tc else "f32",
            "data": "synthetic",
            "config": {"workload": "iVAE 2D rot+trans, 28x28 Bernoulli, latent_dim=2, fc enc / "
                                   "spatial fc dec, batch=512 per GPU (BASELINE configs[1])",
                       "global_batch": BATCH * world, "parallelism": "dp{}".format(world),
                       "decoder_path": "tcgen05-fused" if use_tc else "fp32-generic",
                       "cuda_graphs": bool(svi.use_graphs),
                       "exchange": ("none (1 GPU)" if world == 1 else
                                    "fused all-reduce + Adam kernel over NVLink peer memory"
                                    if svi.peer is not None else "NCCL all-reduce"),
                       "l2": "inputs rotate through a pool of {} batches ({} MB) > 126 MB L2"
                             .format(POOL, POOL * BATCH * H * W * 4 // 2 ** 20)},
            "e2e": {"value": total / t_e2e, "unit": "samples/s",
                    "h2d_bytes_per_step": BATCH * H * W * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches_per_step) * args.steps,
            "launches_per_step": int(launches_per_step),
            "roofline": roof, "kernels": kernels,
            "cpu_baseline": {"value": cpu["value"], "unit": "samples/s", "cores": cpu["cores"],
                             "kind": "port",
                             "sample": "{} steps of batch 512 in {:.1f} s (oracle/svi_port.py)"
                                       .format(cpu["steps"], cpu["seconds"])},
            "clocks": clk.summary(),
            "loss_per_sample": loss_last, "e2e_last_loss_per_sample": loss_f / BATCH / world,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

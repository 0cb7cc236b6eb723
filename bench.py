#!/usr/bin/env python
"""THAT train-step benchmark (BASELINE.json metric: train samples/s; roofline of the dominant kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = augmentation + forward + BCE + backward + Adam over one synthetic batch [B, 3000, F] (train.py:84-101
semantics, dropout and augmentation ON).  `value` times K steps with the batch resident in HBM; `e2e` times the
same K steps through the public API with the batch starting in pinned HOST memory (H2D copy of x and y and a D2H
read of the loss inside the timed region, double-buffered).  N > 1: launched by torchrun, one rank per GPU, batch
sharded data-parallel (B per GPU fixed -> weak scaling), gradients all-reduced with NCCL.
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

T_LEN = 3000
TRAIN_FLOP = {270: 4.90738e9, 540: 16.9008e9}          # per sample, BASELINE.md section 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="samples per GPU")
    ap.add_argument("--features", type=int, default=270, help="270 = one band (config 2), 540 = dual band (config 3)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-op device-time table (stderr)")
    return ap.parse_args()


def out_dim(F):
    return 54 if F == 270 else 90          # activity (6x9) | identity+location+activity (6+30+54)


def synth_batch(B, F, out, seed):
    """x = 20*U[0,1) with a zeroed front pad of U{0..300} rows; y one-hot activity per user."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T_LEN, F, generator=g) * 20
    pad = torch.randint(0, 301, (B,), generator=g)
    for i in range(B):
        x[i, :int(pad[i])] = 0
    y = torch.zeros(B, out)
    users = 6
    per = out // users
    present = torch.rand(B, users, generator=g) > 0.6
    cls = torch.randint(0, max(per, 1), (B, users), generator=g)
    for u in range(users):
        y[torch.arange(B), u * per + cls[:, u]] = present[:, u].float()
    return x, y


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(B, F, out, steps, warmup, threads=None):
    """Times the oracle port of the reference train step (augmentation + fwd + loss + bwd + Adam, dropout on) on the
    host cores.  Returns (samples_per_s, cores, ms_per_step)."""
    import torch
    import torch.nn.functional as Fn
    from oracle import that_oracle as O
    from multi_modal_csi_b200 import THAT
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.set_float32_matmul_precision("highest")
    torch.manual_seed(39)
    sd = {k: v.detach().clone() for k, v in THAT((T_LEN, F), (out,)).state_dict().items()}
    x, y = synth_batch(B, F, out, 1234)
    opt = {}
    names = O.trainable_names(sd)
    drop = lambda t, p: Fn.dropout(t, p, True)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        xa = O.apply_augmentation(x)
        leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
        work = dict(sd)
        work.update(leaves)
        logits = O.that_forward(work, xa, training=True, drop=drop)
        loss = O.bce_with_logits(logits, y)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
        opt["step"] = opt.get("step", 0) + 1
        with torch.no_grad():
            for k, g in zip(names, grads):
                if k not in opt:
                    opt[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
                O.adam_update(sd[k], g, opt[k][0], opt[k][1], opt["step"], 5e-4, 2e-4)
        float(loss.detach())
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return B / (ms / 1e3), threads, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F, out, B = args.features, out_dim(args.features), args.cpu_batch
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    sps, cores, ms = cpu_reference_steps(B, F, out, steps, warm)
    line = {
        "impl": "reference", "metric": "THAT train samples/sec", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"THAT train step, synthetic CSI [B={B},3000,{F}], out={out}, CPU fp32 (reference "
                               f"algorithm via oracle port; BASELINE config 1)", "batch": B},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} train steps of B={B} (augment+fwd+BCE+bwd+Adam, dropout on), torch CPU fp32"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in o.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from multi_modal_csi_b200 import THAT, FusedAdam
    from multi_modal_csi_b200.parallel import GradSync

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F, out, B = args.features, out_dim(args.features), args.batch
    K, W = args.steps, max(args.warmup, 3)

    torch.manual_seed(39)
    model = THAT((T_LEN, F), (out,), act_dtype=args.dtype, max_batch=B).to(dev)
    model.rng_seed = 1000 + rank                       # decorrelated augmentation / dropout per rank
    model.train()
    opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(model, world) if world > 1 else None
    hook = sync                                        # GradSync: bucket 1 is reduced while left encoder 0 runs its backward

    # two distinct resident batches (each 3.3 MB/sample: far larger than the 126 MB L2 at B=256)
    host = [synth_batch(B, F, out, 1234 + rank + 100 * i) for i in range(2)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    def step_resident(i, use_graph=None):
        x, y = resident[i % 2]
        return model.fused_train_step(x, y, opt, pos_weight=4.0, augment=True, grad_hook=hook, use_graph=use_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        step_resident(i)
    barrier()
    eng = model._engine
    ops = eng.ops

    # per-op device time of one step (untimed) -> pick the dominant kernel family for the live roofline
    table = {}
    for _ in range(2):                                 # the first eager pass still pays one-time costs; keep the second
        ops.start_profile()
        step_resident(0, use_graph=False)
        torch.cuda.synchronize(dev)
        table = ops.stop_profile()
    dominant = max(table, key=lambda k: table[k][0]) if table else None
    if args.profile_ops and rank == 0:
        tot = sum(v[0] for v in table.values())
        for k, v in sorted(table.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:24s} {v[0]:9.3f} ms {100 * v[0] / tot:5.1f}%  calls {v[1]}", file=sys.stderr)
        print(f"  total {tot:.3f} ms", file=sys.stderr)

    # ---- value: K steps, batch resident in HBM (forward+loss+backward replayed as one CUDA graph per step)
    launches0 = ops.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)
    clk.__enter__()
    barrier()
    ev0.record()
    for i in range(K):
        step_resident(i)
    ev1.record()
    barrier()
    gpu_launches = ops.launches - launches0
    # ---- roofline: the same K steps launched eagerly with a CUDA-event pair around every launch of the dominant
    #      kernel family (events cannot be placed inside a replayed graph)
    ops.start_profile(only={dominant})
    for i in range(K):
        step_resident(i, use_graph=False)
    dom = ops.stop_profile()
    ms = ev0.elapsed_time(ev1) / K
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B / (ms / 1e3)

    # ---- e2e: same steps from pinned host memory through the public API, H2D/D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(dev)
        stage = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                stage[s][0].copy_(host[s][0], non_blocking=True)
                stage[s][1].copy_(host[s][1], non_blocking=True)
                ready[s].record(copy_stream)

        def run_e2e(n):
            for s in range(2):
                freed[s].record()
            prefetch(0)
            last = None
            for i in range(n):
                s = i % 2
                if i + 1 < n:
                    prefetch(i + 1)
                torch.cuda.current_stream(dev).wait_event(ready[s])
                loss, _ = model.fused_train_step(stage[s][0], stage[s][1], opt, pos_weight=4.0, augment=True,
                                                 grad_hook=hook)
                freed[s].record()
                last = float(loss.item())              # D2H read of the step's result
            return last

        run_e2e(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(K)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1) / K
        t = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        e2e = {"value": world * B / (ems / 1e3), "unit": "samples/s",
               "h2d_bytes_per_step": int(host[0][0].numel() * 4 + host[0][1].numel() * 4), "d2h_bytes_per_step": 4,
               "ms_per_step": ems}

    clk.__exit__()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roof = roofline(dominant, dom, eng, B, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sps, cores, cms = cpu_reference_steps(args.cpu_batch, F, out, 3, 1)
            cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"3 train steps of B={args.cpu_batch} [3000,{F}] (augment+fwd+BCE+bwd+Adam, dropout on), "
                             f"oracle port of the reference on torch CPU fp32, {cms:.0f} ms/step"}
        line = {
            "metric": "THAT train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"THAT train step (augment+fwd+BCE+bwd+Adam, dropout on), synthetic CSI "
                                   f"[B={B}/GPU,3000,{F}], out={out}, {args.dtype} contractions / fp32 master weights",
                       "batch_per_gpu": B, "global_batch": B * world, "features": F, "parallelism": f"dp{world}",
                       "l2": "inputs (829 MB/batch at F=270) larger than the 126 MB L2; two batches alternate"},
            "roofline": roof, "families": family_table(table, peaks), "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": gpu_launches,
            "clocks": clk.summary(),
            "tensor_frac_step": (TRAIN_FLOP.get(F, 0) * value / world) / (peaks.get("bf16_tflops_sustained", 1340.8) * 1e12),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


KERNEL_OF_OP = {"gemm_nt": "gemm_nt_tc3_kernel", "gemm_tn": "gemm_tn_tc3_kernel", "attn_fwd": "attn_fwd_mma_kernel",
                "attn_bwd": "attn_bwd_mma_kernel", "pool_dual": "pool_dual_kernel", "layernorm_bwd": "ln_bwd2_kernel"}


def dram_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel family, from the committed ncu summary."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernel_metrics_step.csv")))
    kern = KERNEL_OF_OP.get(name)
    if not files or not kern:
        return None
    tot = n = 0.0
    for row in csv.DictReader(open(files[-1])):
        if row["kernel"].startswith(kern):
            k = float(row["launches_in_window"])
            tot += k * (float(row["dram__bytes_read.sum"]) + float(row["dram__bytes_write.sum"]))
            n += k
    return tot / n if n else None


def family_table(table, peaks):
    """Every kernel family of one eagerly launched, sequential step (CUDA events around each launch, untimed pass):
    device time, launches and the achieved algorithmic rate against the measured peak that bounds the family."""
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {}
    for name, (ms, calls, work) in sorted(table.items(), key=lambda kv: -kv[1][0]):
        row = {"ms": round(ms, 4), "launches": calls}
        if ms > 0 and work.get("flops", 0) > 0 and name in ("gemm_nt", "gemm_tn", "attn_fwd", "attn_bwd"):
            ach = work["flops"] / (ms * 1e-3) / 1e12
            row.update({"bound": "tensor", "achieved": round(ach, 1), "unit": "TFLOP/s", "frac": round(ach / tf, 3)})
        elif ms > 0 and work.get("bytes", 0) > 0:
            ach = work["bytes"] / (ms * 1e-3) / 1e9
            row.update({"bound": "hbm", "achieved": round(ach, 1), "unit": "GB/s", "frac": round(ach / hbm, 3)})
        out[name] = row
    return out


def roofline(name, dom, eng, B, peaks):
    """Live roofline of the dominant kernel family: algorithmic work per launch / mean launch time."""
    if not name or name not in dom:
        return None
    total_ms, calls, work = dom[name]
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf = peaks.get("bf16_tflops_sustained", 1400.0)
    src = "measured" if peaks else "fallback"
    if work.get("flops", 0) > 0 and name in ("gemm_nt", "gemm_tn", "attn_fwd", "attn_bwd"):
        ach = work["flops"] / (total_ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
                "traffic": dram_traffic(name), "launches": calls, "ms_per_launch": total_ms / calls, "peak_source": src + " (sustained)"}
    ach = work.get("bytes", 0) / (total_ms * 1e-3) / 1e9
    return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "traffic": dram_traffic(name), "launches": calls, "ms_per_launch": total_ms / calls, "peak_source": src}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

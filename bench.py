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
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not bind each rank to its GPU's NUMA node (A/B runs)")
    ap.add_argument("--no-config4", action="store_true", help="skip the CSI-as-image (CNN_2D) line")
    ap.add_argument("--cnn-batch", type=int, default=128, help="samples per GPU of the CSI-as-image line (BASELINE config 4: 128)")
    ap.add_argument("--config3", action="store_true", help="also measure BASELINE config 3 (F=540, out=90) at N=1 (always on for N>1)")
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
REF_STAGED = os.path.join(ROOT, "oracle", "_ref")      # unmodified reference files staged by __graft_entry__.build()


def reference_staged():
    return os.path.isfile(os.path.join(REF_STAGED, "benchmark", "wifi_csi", "model", "that.py"))


def cpu_reference_train(B, F, out, steps, warmup, threads=None):
    """kind "reference": the UNMODIFIED reference (oracle/_ref: model/that.py THAT + train.py train(), staged at build
    time) through its own public API on the host cores: ``train(model, Adam, BCEWithLogitsLoss(pos_weight=4), ...)`` for
    one epoch of `steps` batches (+ the one the loop skips, train.py:81-82) after a warm-up call of `warmup` batches.
    The epoch's own evaluation runs on an 8-sample validation set (train.py:111-127) and is part of the timed call.
    Returns (samples_per_s, cores, ms_per_step)."""
    import contextlib
    import torch
    os.environ.setdefault("CSI_REFERENCE_ROOT", REF_STAGED)
    os.environ.setdefault("WANDB_MODE", "disabled")
    from oracle.ref_import import load_reference
    ref = load_reference()
    try:
        import wandb
        wandb.init(mode="disabled")
    except Exception:
        pass
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(39)
    model = ref.that.THAT((T_LEN, F), (out,))
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=2e-4)          # that.py:395-397
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * out))       # that.py:401
    from torch.utils.data import TensorDataset

    def run(n):
        xs, ys = zip(*[synth_batch(B, F, out, 1234 + i) for i in range(n + 1)])
        train_set = TensorDataset(torch.cat(xs), torch.cat(ys).reshape(-1, 6, out // 6))
        xv, yv = synth_batch(8, F, out, 99)
        valid_set = TensorDataset(xv, yv.reshape(-1, 6, out // 6))
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            try:
                ref.train.train(model=model, optimizer=opt, loss=loss, data_train_set=train_set, data_test_set=valid_set,
                                var_threshold=0.5, var_batch_size=B, var_epochs=1, device=torch.device("cpu"), var_mode="baseline")
            except UnboundLocalError:
                # train.py:175 prints `var_epoch_saved`, which is only bound once an epoch improves on f1 AND PPP: with
                # one epoch on random data the reference raises here, AFTER all of the epoch's work has been done
                pass
        return time.perf_counter() - t0

    run(warmup)
    dt = run(steps)
    ms = 1e3 * dt / steps
    return B / (ms / 1e3), threads, ms


def cpu_reference_steps(B, F, out, steps, warmup, threads=None):
    """kind "port": the oracle restatement of the reference train step (augmentation + fwd + loss + bwd + Adam, dropout
    on) on the host cores -- used when the staged reference files are absent.  Returns (samples_per_s, cores, ms_per_step)."""
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


def cpu_arm(B, F, out, steps, warmup):
    """-> (samples/s, cores, ms/step, kind, sample description)"""
    if reference_staged():
        try:
            sps, cores, ms = cpu_reference_train(B, F, out, steps, warmup)
            return sps, cores, ms, "reference", (
                f"unmodified reference THAT + train() (oracle/_ref), one epoch of {steps} steps of B={B} [3000,{F}] "
                f"(augment+fwd+BCE+bwd+Adam, dropout on; includes the epoch's 8-sample evaluation), torch CPU fp32, {ms:.0f} ms/step")
        except Exception as e:                                       # a broken staging must not lose the baseline
            print(f"[bench] staged reference failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    sps, cores, ms = cpu_reference_steps(B, F, out, steps, warmup)
    return sps, cores, ms, "port", (f"{steps} train steps of B={B} [3000,{F}] (augment+fwd+BCE+bwd+Adam, dropout on), "
                                    f"oracle port of the reference on torch CPU fp32, {ms:.0f} ms/step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F, out, B = args.features, out_dim(args.features), args.cpu_batch
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    sps, cores, ms, kind, sample = cpu_arm(B, F, out, steps, warm)
    line = {
        "impl": "reference", "metric": "THAT train samples/sec", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"THAT train step, synthetic CSI [B={B},3000,{F}], out={out}, CPU fp32 (the reference's own "
                               f"implementation on the host cores; BASELINE config 1)", "batch": B},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
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
def measure(F, out, B, K, W, dtype, dev, world, rank, full, no_e2e=False, profile_ops=False):
    """One configuration on this rank's GPU: `value` (batches resident in HBM, train body replayed as a CUDA graph), the
    live roofline of the dominant kernel family (full only) and `e2e` through the product loader (host -> device inside
    the timed region).  Returns a dict; timings are the max over ranks."""
    import torch
    import torch.distributed as dist
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200 import THAT, FusedAdam
    from multi_modal_csi_b200.loader import CSIBatchSource
    from multi_modal_csi_b200.parallel import GradSync

    torch.manual_seed(39)
    model = THAT((T_LEN, F), (out,), act_dtype=dtype, max_batch=B).to(dev)
    model.rng_seed = 1000 + rank                       # decorrelated augmentation / dropout per rank
    model.train()
    opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(model, world) if world > 1 else None
    hook = sync                                        # GradSync: bucket 1 is reduced while left encoder 0 runs its backward

    # two distinct batches (each 3.3 MB/sample: far larger than the 126 MB L2 at B=256), as ONE host dataset of 2B samples
    hx, hy = zip(*[synth_batch(B, F, out, 1234 + rank + 100 * i) for i in range(2)])
    host_x, host_y = torch.cat(hx), torch.cat(hy)
    resident = [(host_x[i * B:(i + 1) * B].to(dev), host_y[i * B:(i + 1) * B].to(dev)) for i in range(2)]

    def step_resident(i, use_graph=None):
        x, y = resident[i % 2]
        return model.fused_train_step(x, y, opt, pos_weight=4.0, augment=True, grad_hook=hook, use_graph=use_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(W):
        step_resident(i)
    barrier()
    eng = model._engine
    ops = eng.ops
    if dtype == "bf16":
        ops.set_strict_tc(True)                        # from here on a bf16 contraction that cannot run on tcgen05 is an error
    ops.dispatch_counts(reset=True)
    res = {"features": F, "out": out, "batch_per_gpu": B}

    table, dominant, dom = {}, None, {}
    if full:
        # per-op device time of one step (untimed) -> pick the dominant kernel family for the live roofline
        for _ in range(2):                             # the first eager pass still pays one-time costs; keep the second
            ops.start_profile()
            step_resident(0, use_graph=False)
            torch.cuda.synchronize(dev)
            table = ops.stop_profile()
        dominant = max(table, key=lambda k: table[k][0]) if table else None
        if profile_ops and rank == 0:
            tot = sum(v[0] for v in table.values())
            for k, v in sorted(table.items(), key=lambda kv: -kv[1][0]):
                print(f"  {k:24s} {v[0]:9.3f} ms {100 * v[0] / tot:5.1f}%  calls {v[1]}", file=sys.stderr)
            print(f"  total {tot:.3f} ms", file=sys.stderr)

    # ---- value: K steps, batch resident in HBM (forward+loss+backward replayed as one CUDA graph per step)
    launches0 = ops.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        step_resident(i)
    ev1.record()
    barrier()
    res["gpu_launches"] = ops.launches - launches0
    ms = max_over_ranks(ev0.elapsed_time(ev1) / K)
    res["ms_per_step"] = ms
    res["value"] = world * B / (ms / 1e3)
    if full and dominant:
        # ---- roofline: the same K steps launched eagerly with a CUDA-event pair around every launch of the dominant
        #      kernel family (events cannot be placed inside a replayed graph)
        ops.start_profile(only={dominant})
        for i in range(K):
            step_resident(i, use_graph=False)
        dom = ops.stop_profile()
    res["table"], res["dominant"], res["dom"] = table, dominant, dom

    if world > 1:
        # the exchange on its own: one all-reduce of the flat gradient arena (what GradSync issues in two buckets)
        g = model.flat_grads
        for _ in range(2):
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a0.record()
        for _ in range(5):
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        a1.record()
        barrier()
        res["allreduce"] = {"bytes": int(g.numel() * 4), "ms": max_over_ranks(a0.elapsed_time(a1) / 5)}

    # ---- e2e: the same steps through the product loader (multi_modal_csi_b200.loader.CSIBatchSource, "stream" mode, the
    #      code train() runs when the dataset does not stay in HBM): every step's samples are copied from page-locked
    #      host memory into a device staging arena on a copy stream (double-buffered) and the loss is read back
    res["e2e"] = None
    if not no_e2e:
        src = CSIBatchSource(TensorDataset(host_x, host_y), dev, B, mode="stream")
        h2d = [0]

        def run_e2e(n):
            last = None
            lists = [range((i % 2) * B, (i % 2 + 1) * B) for i in range(n)]
            for batch in src.batches(lists):
                loss, _ = model.fused_train_step(batch.x, batch.y, opt, pos_weight=4.0, augment=True, grad_hook=hook,
                                                 offs=batch.offs, lens=batch.lens)
                h2d[0] = batch.h2d_bytes
                last = float(loss.item())              # D2H read of the step's result
            return last

        run_e2e(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(K)
        e1.record()
        barrier()
        ems = max_over_ranks(e0.elapsed_time(e1) / K)
        res["e2e"] = {"value": world * B / (ems / 1e3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d[0]),
                      "d2h_bytes_per_step": 4, "ms_per_step": ems, "loader": "CSIBatchSource(mode='stream')"}
        src.close()
        if full:
            # the same steps fed from the RAGGED host arena (load_data.load_data_x_packed layout: recordings without their
            # zero front pad, which csi_pool_dual re-creates): fewer bytes cross PCIe.  Reported beside `e2e`, which keeps
            # the reference's host format (front-padded dense tensors, load_data.py:66-72).
            from multi_modal_csi_b200.loader import PackedCSIDataset
            lens = [int((host_x[i].abs().sum(dim=1) > 0).nonzero()[0]) for i in range(host_x.shape[0])]      # first non-pad row
            lens = [T_LEN - p for p in lens]
            arena = torch.cat([host_x[i, T_LEN - n:].reshape(-1) for i, n in enumerate(lens)])
            offs, pos = [], 0
            for n in lens:
                offs.append(pos)
                pos += n * F
            src = CSIBatchSource(PackedCSIDataset(arena, offs, lens, F, host_y, T_LEN), dev, B, mode="stream")
            run_e2e(3)
            barrier()
            e0.record()
            run_e2e(K)
            e1.record()
            barrier()
            rms = max_over_ranks(e0.elapsed_time(e1) / K)
            res["e2e_ragged"] = {"value": world * B / (rms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d[0]),
                                 "d2h_bytes_per_step": 4, "ms_per_step": rms,
                                 "loader": "CSIBatchSource(PackedCSIDataset, mode='stream'): unpadded recordings, front pad applied by csi_pool_dual"}
            src.close()
    res["dispatch"] = ops.dispatch_counts()
    res["engine"] = eng
    ops.set_strict_tc(False)
    return res


def measure_cnn2d(B, K, W, dtype, dev, world, rank, no_e2e=False):
    """BASELINE config 4, the CSI-as-image model (the in-reference CNN_2D, model/cnn_2d.py) on [B,3000,270] images through the
    same loader / loss / Adam / data-parallel layers: `value` with the batches resident in HBM, `e2e` through
    CSIBatchSource(stream) with the loss read back every step."""
    import torch
    import torch.distributed as dist
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200 import CNN_2D, FusedAdam
    from multi_modal_csi_b200.loader import CSIBatchSource
    from multi_modal_csi_b200.parallel import GradSync
    F, out = 270, 54
    torch.manual_seed(39)
    model = CNN_2D((T_LEN, F), (out,), act_dtype=dtype, max_batch=B).to(dev)
    model.rng_seed = 2000 + rank
    model.train()
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)            # cnn_2d.py:162-164
    hook = GradSync(model, world) if world > 1 else None
    hx, hy = zip(*[synth_batch(B, F, out, 4321 + rank + 100 * i) for i in range(2)])
    host_x, host_y = torch.cat(hx), torch.cat(hy)
    resident = [(host_x[i * B:(i + 1) * B].to(dev), host_y[i * B:(i + 1) * B].to(dev)) for i in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(i):
        x, y = resident[i % 2]
        return model.fused_train_step(x, y, opt, pos_weight=6.0, augment=True, grad_hook=hook)

    for i in range(W):
        step(i)
    barrier()
    ops = model._engine.ops
    if dtype == "bf16":
        ops.set_strict_tc(True)
    ops.dispatch_counts(reset=True)
    ops.start_profile()
    step(0)
    table = ops.stop_profile()
    launches0 = ops.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        step(i)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1) / K)
    res = {"workload": f"CNN_2D (model/cnn_2d.py) train step on CSI-as-image [B={B}/GPU,3000,{F}], out={out}, {dtype} "
                       f"(BASELINE config 4; the reference has no ResNet-18: CNN_2D is its CSI-as-image model)",
           "value": world * B / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "gpu_launches": (ops.launches - launches0),
           "families": family_table(table, PEAKS()), "e2e": None}
    if not no_e2e:
        src = CSIBatchSource(TensorDataset(host_x, host_y), dev, B, mode="stream")
        h2d = [0]

        def run_e2e(n):
            for batch in src.batches([range((i % 2) * B, (i % 2 + 1) * B) for i in range(n)]):
                loss, _ = model.fused_train_step(batch.x, batch.y, opt, pos_weight=6.0, augment=True, grad_hook=hook,
                                                 offs=batch.offs, lens=batch.lens)
                h2d[0] = batch.h2d_bytes
                float(loss.item())

        run_e2e(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(K)
        e1.record()
        barrier()
        ems = max_over_ranks(e0.elapsed_time(e1) / K)
        res["e2e"] = {"value": world * B / (ems / 1e3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d[0]),
                      "d2h_bytes_per_step": 4, "ms_per_step": ems}
        src.close()
    res["dispatch"] = ops.dispatch_counts()
    ops.set_strict_tc(False)
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and not args.no_numa_bind:
        from multi_modal_csi_b200.parallel import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(dev)                # before any pinned host memory exists (what train() does too)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F, out, B = args.features, out_dim(args.features), args.batch
    K, W = args.steps, max(args.warmup, 3)

    clk = ClockSampler(local)
    clk.__enter__()
    main = measure(F, out, B, K, W, args.dtype, dev, world, rank, full=True, no_e2e=args.no_e2e, profile_ops=args.profile_ops)
    clk.__exit__()
    main.pop("engine")                                 # (its ~1.5 GB of activation buffers are freed before the next configuration)
    # BASELINE config 3 (dual band: 540 features, identity+location+activity = 90 outputs, B=256 per GPU) is quoted at
    # 2/4/8 GPUs: measured in the same launch and reported under "config3" so that the headline config stays config 2
    cfg3 = None
    if (world > 1 or args.config3) and F != 540:
        torch.cuda.empty_cache()
        c3 = measure(540, 90, B, K, W, args.dtype, dev, world, rank, full=False, no_e2e=args.no_e2e)
        c3.pop("engine")
        cfg3 = {"workload": f"THAT dual band [B={B}/GPU,3000,540], out=90 (BASELINE config 3)", "value": c3["value"],
                "unit": "samples/s", "ms_per_step": c3["ms_per_step"], "e2e": c3["e2e"], "allreduce": c3.get("allreduce"),
                "gpu_launches": c3["gpu_launches"], "dispatch": c3["dispatch"],
                "tensor_frac_step": TRAIN_FLOP[540] * c3["value"] / world / 1e12 / None_or(PEAKS().get("bf16_tflops"), 1595.7)}
    cfg4 = None
    if not args.no_config4:
        torch.cuda.empty_cache()
        cfg4 = measure_cnn2d(args.cnn_batch, max(3, K // 2), 3, args.dtype, dev, world, rank, no_e2e=args.no_e2e)
    if rank == 0:
        peaks = PEAKS()
        roof = roofline(main["dominant"], main["dom"], B, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sps, cores, cms, kind, sample = cpu_arm(args.cpu_batch, F, out, 3, 1)
            cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample}
        burst = None_or(peaks.get("bf16_tflops"), 1595.7)
        step_tf = TRAIN_FLOP.get(F, 0) * main["value"] / world / 1e12
        line = {
            "metric": "THAT train samples/sec", "value": main["value"], "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"THAT train step (augment+fwd+BCE+bwd+Adam, dropout on), synthetic CSI "
                                   f"[B={B}/GPU,3000,{F}], out={out}, {args.dtype} contractions / fp32 master weights",
                       "batch_per_gpu": B, "global_batch": B * world, "features": F, "parallelism": f"dp{world}",
                       "l2": "inputs (829 MB/batch at F=270) larger than the 126 MB L2; two batches alternate"},
            "roofline": roof, "families": family_table(main["table"], peaks), "cpu_baseline": cpu, "e2e": main["e2e"],
            "e2e_ragged": main.get("e2e_ragged"),
            "gpu_launches": main["gpu_launches"], "dispatch": main["dispatch"], "allreduce": main.get("allreduce"),
            "clocks": clk.summary(),
            # whole step against the tensor roofline: BASELINE.md section 3 formula, burst cuBLAS peak (sub-second timed region at full clocks)
            "tensor_frac_step": step_tf / burst,
            "tensor_frac_step_sustained": step_tf / None_or(peaks.get("bf16_tflops_sustained"), 1340.8),
            "config3": cfg3, "config4": cfg4, "numa_binding_rank0": numa,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def None_or(v, default):
    return default if v is None else v


_PEAKS = None


def PEAKS():
    global _PEAKS
    if _PEAKS is None:
        try:
            _PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            _PEAKS = {}
    return _PEAKS


KERNEL_OF_OP = {"gemm_nt": "gemm_nt_tc3_kernel", "gemm_tn": "gemm_tn_tc3_kernel", "attn_fwd": "attn_fwd_",
                "attn_bwd": "attn_bwd_", "pool_dual": "pool_dual_kernel", "layernorm_bwd": "ln_bwd2_kernel"}


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
    tf = peaks.get("bf16_tflops", 1595.7)           # burst cuBLAS peak: each launch is timed on its own, at full clocks
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


def roofline(name, dom, B, peaks):
    """Live roofline of the dominant kernel family: algorithmic work per launch / mean launch time.  Tensor-bound
    families are quoted against the BURST cuBLAS bf16 peak of MEASURED_PEAKS.json (the timed region is a fraction of a
    second at full clocks; the sustained figure was measured power-capped at 1237 MHz) -- `frac_sustained` is the same
    number against the sustained peak."""
    if not name or name not in dom:
        return None
    total_ms, calls, work = dom[name]
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf, tfs = peaks.get("bf16_tflops", 1595.7), peaks.get("bf16_tflops_sustained", 1340.8)
    src = "MEASURED_PEAKS.json" if peaks else "fallback"
    if work.get("flops", 0) > 0 and name in ("gemm_nt", "gemm_tn", "attn_fwd", "attn_bwd"):
        ach = work["flops"] / (total_ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
                "frac_sustained": ach / tfs, "traffic": dram_traffic(name), "launches": calls,
                "ms_per_launch": total_ms / calls, "peak_source": src + " bf16_tflops (burst)",
                "flops": "algorithmic (true d, L, k; no head/channel/halo padding), SURVEY 8d"}
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

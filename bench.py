#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task's bench section).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision fp32|bf16]

A "step" is one supervised autoencoder training step (NB:2676-2684: forward, alpha*MSE + CE, backward,
Adam) over one batch of synthetic 3x64x64 images: BASELINE.json configs[1] (batch 256, fp32, 1 B200).
With N > 1 GPUs every rank runs the same per-rank batch (weak scaling) and the flat gradient buffer is
sum-allreduced once per step inside the captured step graph.

value : whole-job images/s with the batches already resident in HBM (rotating over more batches than fit
        in L2), device-timed with CUDA events, max over ranks.
e2e   : the same metric through the public API with HOST (pinned) batches: H2D copy of every batch and a
        D2H read of every step's loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ALPHA, LR = 35.0, 5e-3          # the reference's best configuration (NB:2618)
FLOP_PER_IMAGE_TRAIN = 181.9e6
METRIC = "train images/sec (64x64x3, AE+MLP step)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        tf_burst=float(d["bf16_tflops"]), source="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed regions run: NVML in-process (a query takes microseconds, so a
    70 ms region still gets dozens of samples), `nvidia-smi` as the fallback (one sample per ~0.1 s)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(v) for v in vis.split(",")] if vis and all(v.strip().isdigit() for v in vis.split(",")) else None
            phys = ids[index] if ids and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            mask = 0
        flags = ["Active" if mask & bit else "Not Active" for _, bit in self.NVML_REASONS]
        self.samples.append([str(sm), str(self.max_sm), "0"] + flags)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                              str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            self.samples.append([v.strip() for v in out.split(",")])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nvml is not None:
                    self.nvml = None        # fall back to nvidia-smi for the rest of the run
            time.sleep(0.01 if self.nvml is not None else 0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synthetic_batches(n_batches, batch, seed, pin):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.rand(batch, 3, 64, 64, generator=g) for _ in range(n_batches)]
    ys = [torch.randint(0, 10, (batch,), generator=g) for _ in range(n_batches)]
    if pin:
        xs = [x.pin_memory() for x in xs]
        ys = [y.pin_memory() for y in ys]
    return xs, ys


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm.  kind "reference": the reference's OWN classes (the notebook's cells, executed by
# oracle/load_reference.py from /root/reference or from the git-ignored copy under baseline/_ref) driven by the literal loop
# body NB:2676-2684 with torch.optim.Adam (NB:2654).  kind "port": oracle/torch_port.py's functional restatement with the
# same torch.optim.Adam, when no notebook can be found.
# --------------------------------------------------------------------------------------------------
def reference_step_fn(device):
    """Returns (step(x, y) -> loss tensor, kind, description): one iteration of NB:2676-2684 on `device` in stock PyTorch."""
    import torch.nn as nn
    try:
        from oracle import load_reference
        if load_reference.reference_available():
            cls = load_reference.load_reference_classes()
            torch.manual_seed(0)
            model = cls["SupervisedAutoencoder"](64, 10).to(device)
            model.train()
            mse, ce = nn.MSELoss(), nn.CrossEntropyLoss()
            opt = torch.optim.Adam(model.parameters(), lr=LR)

            def step(x, y):
                opt.zero_grad()
                x_hat, logits, _ = model(x)
                loss = ALPHA * mse(x_hat, x) + ce(logits, y)
                loss.backward()
                opt.step()
                return loss
            return step, "reference", "the notebook's SupervisedAutoencoder + loop body NB:2676-2684, torch.optim.Adam"
    except Exception as ex:                                   # fall through to the port
        print(f"bench: reference classes unavailable ({ex}); timing the oracle port", file=sys.stderr)
    from oracle import seeded, torch_port as tp
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), 0)
    st = {k: v.to(device) for k, v in st.items()}
    keys = tp.param_keys(st)
    leaves = {k: st[k].detach().clone().requires_grad_(True) for k in keys}
    opt = torch.optim.Adam([leaves[k] for k in keys], lr=LR)

    def step(x, y):
        opt.zero_grad()
        work = dict(st)
        work.update(leaves)
        nb = {}
        x_hat, logits, _ = tp.ae_forward(work, x, True, nb)
        loss, _, _ = tp.ae_loss(x_hat, logits, x, y, ALPHA)
        loss.backward()
        opt.step()
        st.update(nb)
        return loss
    return step, "port", "oracle/torch_port.py forward + torch autograd + torch.optim.Adam"


def cpu_train_steps(batch, steps, warmup):
    torch.manual_seed(0)
    step, kind, what = reference_step_fn(torch.device("cpu"))
    x = torch.rand(batch, 3, 64, 64)
    y = torch.randint(0, 10, (batch,))
    for _ in range(warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, y).item()                                     # NB:2687: loss.item() every step
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps = max(1, min(args.steps, 40))
    ips, per, kind, what = cpu_train_steps(args.batch, steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "supervised AE train step (alpha*MSE + CE, Adam), batch 256, 3x64x64, latent 64",
                   "batch_per_step": args.batch},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{steps} train steps of batch {args.batch} ({what}; torch CPU fp32)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# Stock PyTorch on the same B200 (SURVEY 8d "the real bar"): the reference modules, eager, default torch settings, the
# literal loop body with a device-resident batch; and the drop-in loop of INTEGRATION.md section 1 (the same Python loop
# with ae_b200's modules and optimizer: three autograd Functions + the fused flat Adam, no whole-step graph).
# --------------------------------------------------------------------------------------------------
def _time_loop(step, xs, ys, steps, warmup):
    for i in range(warmup):
        step(xs[i % len(xs)], ys[i % len(ys)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(steps):
        last = step(xs[i % len(xs)], ys[i % len(ys)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps, float(last.detach())


def torch_gpu_baseline(dev, B, xs, ys, steps=50, warmup=10):
    step, kind, what = reference_step_fn(dev)
    t, loss = _time_loop(step, xs, ys, steps, warmup)
    return {"value": B / t, "unit": "images/s", "ms_per_step": t * 1e3, "kind": kind, "what": what + "; eager, CUDA, device-resident batches",
            "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32), "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
            "steps": steps, "final_loss": loss}


def dropin_loop(dev, B, xs, ys, precision, backend, steps=50, warmup=10):
    import torch.nn as nn
    import ae_b200
    torch.manual_seed(0)
    model = ae_b200.SupervisedAutoencoder(64, 10, precision=precision, backend=backend).to(dev).train()
    model.engine().prepare(dev, B)
    opt = ae_b200.Adam(model.parameters(), lr=LR)
    mse, ce = nn.MSELoss(), nn.CrossEntropyLoss()

    def step(x, y):                                           # INTEGRATION.md section 1 = NB:2676-2684 verbatim
        opt.zero_grad()
        x_hat, logits, _ = model(x)
        loss = ALPHA * mse(x_hat, x) + ce(logits, y)
        loss.backward()
        opt.step()
        return loss
    t, loss = _time_loop(step, xs, ys, steps, warmup)
    return {"value": B / t, "unit": "images/s", "ms_per_step": t * 1e3, "steps": steps, "final_loss": loss,
            "what": "model(imgs); alpha*mse + ce; loss.backward(); optimizer.step() with ae_b200 modules (no step graph)"}


def full_pipeline(dev):
    """BASELINE configs[2]: AE pre-training, latent extraction and the MLP on frozen latents over a synthetic, class-structured
    27,000-image EuroSAT-shaped set (18,900 / 4,050 / 4,050), bf16 mode, everything on the device (scripts/full_pipeline.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ae_full_pipeline", os.path.join(ROOT, "scripts", "full_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.run(precision="bf16", dev=dev)


def mlp_train_rate(dev, steps=200, batch=64):
    """Second stage (NB:3475-3486): MLP steps on frozen 64-d latents, vectors/s, through the public fit helper's step."""
    import ae_b200
    from ae_b200 import fit
    torch.manual_seed(0)
    clf = ae_b200.MLP(64, 10).to(dev).train()
    clf._state.prepare(dev, batch)
    opt = ae_b200.Adam(clf.parameters(), lr=1e-4, weight_decay=1e-4)
    n = batch * 64
    X, y = torch.randn(n, 64, device=dev), torch.randint(0, 10, (n,), device=dev)
    fit.train_epoch_mlp(clf, opt, X, y, batch, True, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(1, steps // 64)
    for _ in range(reps):
        fit.train_epoch_mlp(clf, opt, X, y, batch, True, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": reps * n / dt, "unit": "vectors/s", "batch": batch, "us_per_step": dt / (reps * 64) * 1e6}


# --------------------------------------------------------------------------------------------------
# Roofline of the dominant kernel: the TMA-fed tcgen05 row GEMM (k_tma_rowgemm, 12 launches per step, the largest
# share of the step in profiles/*launches*).  Timed on its own with CUDA events on the launching stream, on the
# geometry of its most frequent instance (ConvTranspose2d 128->64 forward, NB:620: 8x8 -> 16x16 pixels), inputs
# rotating over more buffers than fit in L2.  Algorithmic bytes per launch: the operand
# planes once + the packed weights once + the fp32 output once.
# --------------------------------------------------------------------------------------------------
def measured_traffic(precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this very kernel from the round's `ncu --set full` capture
    (profiles/r2_roofline_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep); None when no capture of the
    current kernel generation has been committed."""
    p = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
    try:
        return json.load(open(p)).get(precision)
    except Exception:
        return None


def kernel_roofline(dev, B, precision, replays=20):
    """The dominant kernel on its own: ConvTranspose2d 128->64 forward (NB:620, 8x8 -> 16x16 pixels) at batch B through the
    C ABI.  The launches are captured ONCE into a CUDA graph (one launch per buffer set, rotating over more sets than fit
    in L2) and the graph is replayed between CUDA events on the launching stream: the average is device time per launch,
    launch gaps included, host-side tensor-map encoding excluded."""
    import ctypes as C
    from ae_b200 import _lib
    lib = _lib.load()
    prec = _lib.PREC_FP32 if precision == "fp32" else _lib.PREC_BF16
    nsplit = 2 if precision == "fp32" else 1
    hs, cb, cs = 8, 64, 128
    M = B * hs * hs
    g = _lib.ConvGeom(B, hs, hs, cb, cs)
    w = torch.randn(cs, cb, 3, 3, device=dev) / (9 * cs / 4) ** 0.5
    nbytes = lib.ae_packed_weight_bytes(cs, cb, prec, _lib.BACKEND_TC)
    raw = torch.zeros(2 * nbytes + 2048, dtype=torch.uint8, device=dev)
    base = (raw.data_ptr() + 1023) & ~1023
    pk_f, pk_d = C.c_void_p(base), C.c_void_p((base + nbytes + 1023) & ~1023)
    _lib.check(lib.ae_pack_conv_weight(_lib.ptr(w), cs, cb, pk_f, pk_d, prec, _lib.BACKEND_TC, _lib.stream_ptr()))
    bias = torch.zeros(cb, device=dev)
    stats = torch.zeros(2 * cb, dtype=torch.float64, device=dev)
    a_bytes, o_bytes, w_bytes = M * cs * 2 * nsplit, 4 * M * cb * 4, 9 * cs * cb * 2 * nsplit
    n_rot = max(2, int(140e6 // (a_bytes + o_bytes)) + 1)
    planes = [(torch.randn(nsplit * M * cs, device=dev) * 0.5).to(torch.bfloat16) for _ in range(n_rot)]
    outs = [torch.empty(B, 2 * hs, 2 * hs, cb, device=dev) for _ in range(n_rot)]
    ep = _lib.Epilogue(_lib.EPI_BIAS_STATS, _lib.ptr(bias), None, None, _lib.ptr(stats))
    st = torch.cuda.Stream(device=dev)

    def launch(i):
        op = _lib.Operand(_lib.ptr(planes[i % n_rot]), None, None, 0.0, _lib.OP_SPLIT_BF16)
        _lib.check(lib.ae_conv2d_s2_dgrad(C.byref(g), C.byref(op), pk_d, C.byref(ep), _lib.ptr(outs[i % n_rot]), prec,
                                          _lib.BACKEND_TC, C.c_void_p(st.cuda_stream)))

    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for i in range(n_rot):
            launch(i)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for i in range(n_rot):
                launch(i)
        for _ in range(3):
            graph.replay()
        st.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(replays)]
        for e0, e1 in evs:
            e0.record(st)
            graph.replay()
            e1.record(st)
        st.synchronize()
    ts = sorted(e0.elapsed_time(e1) * 1e-3 / n_rot for e0, e1 in evs)
    t = sum(ts) / len(ts)
    peaks = measured_peaks()
    alg = a_bytes + o_bytes + w_bytes
    flop = 2.0 * M * 9 * cs * cb
    return {"bound": "hbm", "achieved": alg / t / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "frac": alg / t / 1e9 / peaks["hbm"],
            "traffic": measured_traffic(precision),
            "kernel": "second-generation row GEMM k_rowgemm2<DGRAD,64> (ConvTranspose2d 128->64 forward, batch %d)" % B,
            "algorithmic_bytes": alg, "avg_launch_us": t * 1e6, "median_launch_us": ts[len(ts) // 2] * 1e6,
            "launches_timed": replays * n_rot, "timing": "CUDA-graph replays of %d back-to-back launches between CUDA events" % n_rot,
            "tensor_tflops": flop / t / 1e12, "tensor_frac_of_burst": flop / t / 1e12 / peaks["tf_burst"],
            "peak_source": peaks["source"], "l2": f"inputs/outputs rotate over {n_rot} buffer sets ({n_rot * (a_bytes + o_bytes) / 1e6:.0f} MB > L2)"}


FLOP_PER_IMAGE_INFER = 30.64e6   # encoder + MLP forward


def inference_rate(dev, precision, backend, batch=4096, iters=20, sweep=(1024, 4096, 16384, 65536)):
    """BASELINE's second metric: encoder + MLP inference images/s (clf(enc(x)).argmax(1), eval mode), device-resident
    inputs rotating over > L2, CUDA events.  `sweep`: BASELINE configs[4] (batch 1k-64k), the same measurement per batch
    size (batches above AE_B200_EVAL_CHUNK are walked in chunks by the encoder shell)."""
    import ae_b200
    torch.manual_seed(1)
    ae = ae_b200.SupervisedAutoencoder(64, 10, precision=precision, backend=backend).to(dev).eval()
    clf = ae_b200.MLP(64, 10).to(dev).eval()
    peaks = measured_peaks()

    def rate(b, n_it):
        n_in = max(2, min(3, int(260e6 // (b * 49152)) + 1))          # > 2 x L2 of distinct input bytes in rotation
        xs = [torch.rand(b, 3, 64, 64, device=dev) for _ in range(n_in)]
        for i in range(3):
            ae_b200.encode_predict(ae.enc, clf, xs[i % n_in])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_it):
            ae_b200.encode_predict(ae.enc, clf, xs[i % n_in])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / n_it

    t = rate(batch, iters)
    out = {"metric": "encoder+MLP infer img/s", "value": batch / t, "unit": "images/s", "batch": batch, "ms_per_batch": t * 1e3,
           "dtype": precision, "tflops": batch / t * FLOP_PER_IMAGE_INFER / 1e12,
           "input_gbs": batch / t * 49152 / 1e9, "frac_of_hbm_input_bound": batch / t * 49152 / 1e9 / peaks["hbm"]}
    if sweep:
        out["sweep"] = []
        for b in sweep:
            tb = rate(b, max(3, min(iters, int(200000 // b) + 1)))
            out["sweep"].append({"batch": b, "images_per_s": b / tb, "ms_per_batch": tb * 1e3})
    return out


# --------------------------------------------------------------------------------------------------
# BASELINE configs[3]: data-parallel training at GLOBAL batch 4096 (4096 / world images per rank), and the
# on-hardware parity check of the data-parallel step (per-rank BatchNorm statistics, gradients averaged).
# --------------------------------------------------------------------------------------------------
def dp_global_batch(dev, world, rank, comm, args, global_batch, steps=30, warmup=5):
    import torch.distributed as dist
    import ae_b200
    b = global_batch // world
    torch.manual_seed(0)
    model = ae_b200.SupervisedAutoencoder(64, 10, precision=args.precision, backend=args.backend).to(dev).train()
    model.engine().prepare(dev, b)
    if world > 1:
        ae_b200.dp.broadcast_parameters(model)
    opt = ae_b200.Adam(model.parameters(), lr=LR)
    stepper = ae_b200.TrainStep(model, opt, ALPHA, b, comm=comm)
    n_rot = max(2, int(300e6 // (b * 49152)) + 1)             # distinct input bytes in rotation > 2 x L2
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    xs = [torch.rand(b, 3, 64, 64, device=dev, generator=g) for _ in range(n_rot)]
    ys = [torch.randint(0, 10, (b,), device=dev, generator=g) for _ in range(n_rot)]

    def run(k, i0):
        for i in range(k):
            stepper.load(xs[(i0 + i) % n_rot], ys[(i0 + i) % n_rot])
            stepper.run()
    run(warmup, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stepper.stream)
    run(steps, warmup)
    e1.record(stepper.stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    loss = float(stepper.loss[0])
    stepper.close()
    return {"workload": f"BASELINE configs[3]: data-parallel AE train step, GLOBAL batch {global_batch} = {b} per GPU x {world}",
            "value": global_batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "batch_per_gpu": b, "n_gpus": world,
            "scaling": "strong", "steps": steps, "kernels_per_step": int(stepper.num_kernels), "final_loss": loss,
            "step_tflops": FLOP_PER_IMAGE_TRAIN * global_batch / (ms * 1e-3) / 1e12}


def dp_gradient_check(dev, world, rank, comm, model, opt, stepper, xs_d, ys_d):
    """One more data-parallel step, checked on rank 0 against the single-GPU path: gather every rank's shard, compute each
    shard's gradient with `train_step_grads` on the same weights, average, and compare with (a) the gradient buffer the
    captured step leaves behind (the NCCL sum) times 1/world and (b) the parameters after the step, predicted by torch's
    Adam formula from the averaged gradient and the optimizer state before the step (this is what a wrong 1/world in the
    fused Adam kernel would break)."""
    import torch.distributed as dist
    import ae_b200
    eng = model.engine()
    flat = eng.flat
    st = opt.flat_state(flat)
    B = xs_d[0].shape[0]
    x, y = xs_d[rank % len(xs_d)], ys_d[rank % len(ys_d)]
    torch.cuda.synchronize()
    p0, m0, v0 = flat.data.clone(), st["m"].clone(), st["v"].clone()
    step0 = int(st["step"][0])
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    stepper.load(x, y)
    stepper.run()
    torch.cuda.synchronize()
    fused = getattr(flat, "_dp_flags", None) is not None      # reduce-scatter + Adam + all-gather in one kernel over peer memory
    g_nccl = flat.grad.clone()
    if fused:                                                 # the gradient buffer then still holds this rank's OWN gradient
        dist.all_reduce(g_nccl)
    g_nccl /= world
    p1 = flat.data.clone()
    gx = [torch.empty_like(x) for _ in range(world)]
    gy = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gx, x)
    dist.all_gather(gy, y)
    g_local = None
    if fused:                                                 # every rank's own gradient of this step, for a per-shard comparison
        g_local = [torch.empty_like(flat.grad) for _ in range(world)]
        dist.all_gather(g_local, flat.grad.clone())
    out = None
    if rank == 0:
        ref = ae_b200.SupervisedAutoencoder(64, 10, precision=eng.precision, backend=eng.backend).to(dev).train()
        acc = torch.zeros(flat.len, dtype=torch.float64, device=dev)
        per_shard, per_shard_noise = [], []
        for r in range(world):
            ref.load_state_dict(state0)
            ref.train_step_grads(gx[r], gy[r], ALPHA)
            gr = ref.engine().flat.grad.double()
            acc += gr
            if g_local is not None:
                per_shard.append(float((g_local[r].double() - gr).norm() / gr.norm()))
            # the same shard once more: run-to-run noise floor of the single-GPU path for THIS data (fp32 shared-memory
            # atomics of the BatchNorm statistics arrive in another order every launch; a pre-activation within an ulp of zero
            # then takes the other ReLU branch, and BatchNorm's backward amplifies that)
            ref.load_state_dict(state0)
            ref.train_step_grads(gx[r], gy[r], ALPHA)
            per_shard_noise.append(float((ref.engine().flat.grad.double() - gr).norm() / gr.norm()))
        g_ref = (acc / world).float()
        grad_rel = float((g_nccl - g_ref).norm() / g_ref.norm())
        noise = max(per_shard_noise)
        group = opt.param_groups[0]
        b1, b2 = group["betas"]
        t = step0 + 1
        m = m0 + (g_ref - m0) * (1 - b1)
        v = v0 * b2 + g_ref * g_ref * (1 - b2)
        denom = v.sqrt() / (1 - b2 ** t) ** 0.5 + group["eps"]
        p_pred = p0 - (group["lr"] / (1 - b1 ** t)) * (m / denom)
        # fused path: every rank keeps Adam's moments of its own 1/world shard only -- rank 0 can predict its shard, the first
        # ceil(len/4 / world) float4s of the flat buffer (dp.cu dp_adam_fused: ONE exchange round over the whole buffer)
        own = slice(0, min(flat.len, ((flat.len // 4 + world - 1) // world) * 4)) if fused else slice(0, flat.len)
        upd_rel = float((p1[own] - p_pred[own]).norm() / (p_pred[own] - p0[own]).norm())
        # the same prediction with the SUM instead of the mean: how far off a missing 1/world would be
        ms_, vs_ = m0 + (g_ref * world - m0) * (1 - b1), v0 * b2 + (g_ref * world) ** 2 * (1 - b2)
        p_bad = p0 - (group["lr"] / (1 - b1 ** t)) * (ms_ / (vs_.sqrt() / (1 - b2 ** t) ** 0.5 + group["eps"]))
        tol = lambda n: max(1e-5, 10.0 * n)
        # the noise is bursty (one pre-activation within an ulp of zero taking the other ReLU branch moves a shard's gradient by
        # 1e-4): the floor is the LARGEST run-to-run difference seen over the shards, not each shard's single sample
        ok = upd_rel <= tol(noise) and grad_rel <= tol(noise) and all(d <= tol(noise) for d in per_shard)
        out = {"ok": bool(ok), "criterion": "every difference <= max(1e-5, 10 x the largest run-to-run difference of the single-GPU path over the shards)",
               "path": "fused reduce-scatter + Adam + all-gather over NVLink peer memory (k_dp_adam)" if fused else "NCCL allreduce + Adam",
               "grad_rel": grad_rel, "grad_rel_run_to_run_single_gpu": noise, "grad_rel_per_rank": per_shard,
               "grad_rel_run_to_run_per_shard": per_shard_noise, "update_rel": upd_rel,
               "update_rel_if_scale_were_missing": float((p1[own] - p_bad[own]).norm() / (p_bad[own] - p0[own]).norm()),
               "world": world, "adam_step": t, "batch_per_gpu": int(B),
               "what": "rel-L2 of NCCL gradient/world (and of the parameters after the captured step) against per-shard single-GPU gradients averaged on rank 0"}
    dist.barrier()
    return out


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import ae_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = ae_b200.dp.init_communicator()
    B = args.batch
    torch.manual_seed(0)
    model = ae_b200.SupervisedAutoencoder(64, 10, precision=args.precision, backend=args.backend).to(dev).train()
    model.engine().prepare(dev, B)
    if world > 1:
        ae_b200.dp.broadcast_parameters(model)          # also invalidates the weight packs: TrainStep re-derives them
    opt = ae_b200.Adam(model.parameters(), lr=LR)
    stepper = ae_b200.TrainStep(model, opt, ALPHA, B, comm=comm)

    n_rot = 12   # 12 x 12.6 MB of inputs > 126 MB of L2: every step reads a batch that is not L2-resident
    xs_h, ys_h = synthetic_batches(n_rot, B, 1234 + rank, pin=True)
    xs_d = [x.to(dev) for x in xs_h]
    ys_d = [y.to(dev) for y in ys_h]
    stream = stepper.stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_steps(k, i0=0):
        for i in range(k):
            stepper.load(xs_d[(i0 + i) % n_rot], ys_d[(i0 + i) % n_rot])
            stepper.run()

    # ---- device-resident throughput ----
    sampler = ClockSampler(local)
    sampler.start()
    device_steps(args.warmup)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    device_steps(args.steps, args.warmup)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)
    final_loss = [float(v) for v in stepper.loss[:3].cpu()]

    # ---- end to end: pinned host batches in, loss out, every step (public API: TrainStep.run_batches) ----
    def host_batches(k, i0=0):
        for i in range(k):
            yield xs_h[(i0 + i) % n_rot], ys_h[(i0 + i) % n_rot]

    stepper.run_batches(host_batches(0 if args.no_e2e else min(args.warmup, 3)))
    barrier()
    w0 = time.perf_counter()
    losses = stepper.run_batches(host_batches(1 if args.no_e2e else args.steps))
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    barrier()
    tt = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = B * world * (1 if args.no_e2e else args.steps) / float(tt.item())
    clocks = sampler.summary()      # sampled from the warm-up through both timed regions (device-resident and end-to-end)

    # ---- the reference's own loop shape (NB:2672-2688): copy, step, `loss.item()` -- one host sync per step, no overlap ----
    e2e_sync = None
    if not args.no_e2e:
        try:
            for i in range(3):
                float(stepper(xs_h[i % n_rot], ys_h[i % n_rot])[0])
            barrier()
            w0 = time.perf_counter()
            for i in range(args.steps):
                float(stepper(xs_h[i % n_rot], ys_h[i % n_rot])[0])          # H2D from pinned memory, graph, D2H of the loss
            wall = time.perf_counter() - w0
            barrier()
            ts = torch.tensor([wall], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            e2e_sync = B * world * args.steps / float(ts.item())
        except Exception as ex:                                              # a reporting extra must never cost the headline
            print(f"bench: synchronous end-to-end leg skipped: {ex}", file=sys.stderr)
    assert len(losses) == (1 if args.no_e2e else args.steps) and all(torch.isfinite(l).all() for l in losses)

    # ---- the dominant kernel on its own stream-ordered CUDA events (roofline), and encoder+MLP inference ----
    roof = None if args.no_roofline else kernel_roofline(dev, B, args.precision)
    infer = None if args.no_roofline else inference_rate(dev, args.precision, args.backend)

    # ---- extras (reported, never the headline): stock PyTorch on this GPU, the drop-in loop, the MLP stage, and the
    # data-parallel configuration BASELINE names (global batch 4096, strong scaling) with an on-hardware gradient check
    extras = {}
    if not args.no_extras:
        def guarded(name, fn):
            try:
                extras[name] = fn()
            except Exception as ex:                          # a reporting extra must never cost the headline
                extras[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        if rank == 0 or world == 1:
            guarded("torch_gpu_baseline", lambda: torch_gpu_baseline(dev, B, xs_d, ys_d))
            guarded("dropin_loop", lambda: dropin_loop(dev, B, xs_d, ys_d, args.precision, args.backend))
            guarded("mlp_train", lambda: mlp_train_rate(dev))
            if args.precision == "fp32" and args.backend == "tc":
                # the same inference metric in the library's bf16 mode (one operand plane, fp32 accumulate; BASELINE's
                # stated tolerance for it is rel <= 1e-2): reported beside the fp32-mode headline, never in its place
                guarded("inference_bf16_mode", lambda: inference_rate(dev, "bf16", args.backend, sweep=()))
            guarded("full_pipeline_bf16", lambda: full_pipeline(dev))
        barrier()
        guarded("dp_global_4096", lambda: dp_global_batch(dev, world, rank, comm, args, 4096))
        if world > 1:
            guarded("dp_check", lambda: dp_gradient_check(dev, world, rank, comm, model, opt, stepper, xs_d, ys_d))
        barrier()

    peaks = measured_peaks()
    flops = FLOP_PER_IMAGE_TRAIN * B
    step_s = ms * 1e-3 / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": "supervised AE train step (alpha*MSE + CE, Adam), batch 256 per GPU, 3x64x64, latent 64",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "precision_mode": ("fp32 storage, tcgen05 2-term bf16 operand split, fp32 accumulate" if args.precision == "fp32"
                                      else "bf16 operands, fp32 accumulate") if args.backend == "tc" else "fp32 CUDA cores",
                   "backend": args.backend, "l2": f"inputs rotate over {n_rot} batches ({n_rot * B * 49152 / 1e6:.0f} MB > L2)",
                   "final_loss": final_loss},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": B * (49152 + 8), "d2h_bytes_per_step": 16,
                "pipelined": "TrainStep.run_batches: H2D of batch i+1 and D2H of loss i-1 overlap step i",
                "sync_every_step": e2e_sync},
        "gpu_launches": int(stepper.num_kernels) * args.steps,
        "kernels_per_step": int(stepper.num_kernels),
        "clocks": clocks,
        "step_tflops": flops / step_s / 1e12,
    }
    if roof is not None:
        line["roofline"] = roof
    if infer is not None:
        line["inference"] = infer
    line.update(extras)
    if rank == 0 and not args.no_cpu_baseline and world == 1:      # N=1 only: other ranks' spin-waits would share the host cores
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = 12
        ips, per, kind, what = cpu_train_steps(B, n, 2)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"{n} train steps of batch {B} ({what}; torch CPU fp32)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        # every rank applied the same averaged gradient to the same start weights: the replicas must be bit-identical
        flat = model.engine().flat.data
        digest = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        in_sync = all(torch.equal(g, gathered[0]) for g in gathered)
        if rank == 0:
            print(json.dumps({"dp_check": "parameters identical on all ranks after the timed steps", "ok": bool(in_sync),
                              "world": world}), file=sys.stderr)
        assert in_sync, "data-parallel replicas diverged"
        # the captured graph holds NCCL kernels of this communicator: release it before the communicator goes away
        stepper.close()
        torch.cuda.synchronize()
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


def main():
    # a hung collective must not hold a multi-GPU box: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("AE_BENCH_WATCHDOG_S", "900")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("AE_B200_PRECISION", "fp32"), choices=["fp32", "bf16"])
    ap.add_argument("--backend", default=os.environ.get("AE_B200_BACKEND", "tc"), choices=["tc", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the stand-alone kernel timing and the inference leg")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the end-to-end leg (its number is then meaningless)")
    ap.add_argument("--roofline-only", action="store_true", help="profiling aid: only the stand-alone launches of the roofline kernel (for ncu --set full)")
    ap.add_argument("--no-extras", action="store_true", help="skip the torch-GPU baseline, drop-in loop, MLP and global-batch-4096 legs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.roofline_only:
        torch.cuda.set_device(0)
        print(json.dumps(kernel_roofline(torch.device("cuda", 0), args.batch, args.precision)))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

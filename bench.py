#!/usr/bin/env python
"""bench.py -- CIDNet inference throughput (megapixels/s) on B200, per the driver contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4|cfg5|cfg3]

With N > 1 and no --workload the headline line is still the cfg2 replica run (weak scaling), and the SAME process
group then also measures the two multi-GPU configs BASELINE.json names -- cfg5 (one 4K image, row strips over the N
GPUs, halos + Gram all-reduce over NCCL, CUDA-graph replay) and cfg4 (64 images split N ways) -- each with an in-run
parity check, reported under `extra_workloads` in the one JSON line.

A "step" is ONE pass of the hot path (CIDNet.forward through the C ABI / sm_100a kernels)
over one batch of synthetic input of the named shape.  Default workload = BASELINE.json
configs[1]: 1x3x640x1120 (LOL-Blur.pth when present, else seeded random-init weights through
a real .pth round trip).  Multi-GPU: one process per GPU (torchrun), images are independent
units -> every rank runs its own batch, no data-path collective ("weak" scaling); the timing
is the max over ranks of the device-timed region.

JSON line keys: see the contract in the task description; extras: roofline (dominant kernel,
timed live with CUDA events recorded on the launching stream around every kernel launch),
kernels (per-kernel share), cpu_baseline (oracle port on host cores, bounded sample), e2e
(pinned host -> H2D -> forward -> D2H every step), clocks, gpu_launches.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, H, W, description)
    "cfg1": (1, 400, 600, "CIDNet 1x3x400x600 (LOLv1 shape)"),
    "cfg2": (1, 640, 1120, "CIDNet 1x3x640x1120 (BASELINE.json configs[1])"),
    "cfg4": (64, 400, 600, "CIDNet 64x3x400x600 batch (per rank: 64/N images)"),
    "cfg5": (1, 2160, 3840, "CIDNet 1x3x2160x3840 single 4K image (N>1: rows sharded over the GPUs)"),
    "cfg3": (32, 1080, 1920, "standalone PHVIT(HVIT(x)) round trip, 32x3x1080x1920 (HBM roofline check)"),
}
METRIC = "CIDNet inference megapixels/s"
UNIT = "MP/s"
L2_BYTES = 126 * 1024 * 1024


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def find_weights():
    for p in (os.environ.get("CIDNET_WEIGHTS"), os.path.join(ROOT, "LOL-Blur.pth"), "/root/reference/LOL-Blur.pth"):
        if p and os.path.exists(p):
            return p
    return None


def make_weights(tmpdir="/tmp", mssa=False):
    """state_dict for the bench: shipped LOL-Blur.pth if present, else seeded random-init saved and
    re-loaded through a real .pth file (same path a user takes: eval_SID_blur.py:22)."""
    import torch
    w = None if mssa else find_weights()
    if w:
        return torch.load(w, map_location="cpu"), "LOL-Blur.pth"
    from oracle.cidnet_oracle import make_state_dict
    sd = make_state_dict(0, perturb=False, mssa=mssa)
    path = os.path.join(tmpdir, f"cidnet_bench_{os.getpid()}.pth")
    torch.save(sd, path)
    sd = torch.load(path, map_location="cpu")
    os.remove(path)
    return sd, "seeded random-init (.pth round trip; LOL-Blur.pth absent)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (a polling
    `nvidia-smi -lms` process measurably disturbs kernel launches; NVML calls from a thread do not)."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # map the CUDA ordinal to NVML through the PCI bus id (CUDA_VISIBLE_DEVICES safe)
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            # warm every query once: the first call of some NVML entry points takes tens of ms (driver RPC) and
            # was seen to stall the running kernels when it happened inside the timed region
            for _ in range(2):
                pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pynvml.nvmlDeviceGetPowerUsage(self.h)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, mx, rs, pw))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.02)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def poll_until(self, event):
        """Sample from the CALLING thread while `event` (a recorded torch.cuda.Event) is pending: the
        work is already enqueued, so the samples are taken under load without a second thread
        competing with the launch loop."""
        if self.nv is None:
            event.synchronize()
            return
        nv = self.nv
        self.thread = True
        while not event.query():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.01)

    def stop(self):
        if self.nv is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag = True
        if self.thread is not True:
            self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max((r[3] for r in self.rows), default=None)}


def run_reference(args):
    """--impl reference: the reference's CPU eval path (demo.py:26-27,55-57) timed on the host cores.
    /root/reference does not exist on the GPU box, so the arm runs the oracle port of the same
    algorithm (oracle/cidnet_oracle.py, pinned to the reference by tests/golden) with all host threads,
    including the dead I_LCA5 the reference executes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cidnet_oracle as O
    O.FAST_BILINEAR = True
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, H, W, desc = WORKLOADS[args.workload if args.workload != "cfg3" else "cfg2"]
    mssa = args.variant == "mssa"
    sd, wdesc = make_weights(mssa=mssa)
    sd = {k: v.float() for k, v in sd.items()}
    sample_B, sample_H, sample_W = 1, H, W
    x = O.make_input("uniform", sample_B, sample_H, sample_W, seed=1234)
    t0 = time.perf_counter(); O.forward(x, sd, run_dead_block=True, mssa=mssa); t1 = time.perf_counter() - t0
    budget = 90.0
    if (args.steps + args.warmup) * t1 > budget:      # bounded sample: a centre crop (multiple of 8)
        f = max(0.1, (budget / ((args.steps + args.warmup) * t1)) ** 0.5)
        sample_H, sample_W = max(64, int(H * f) // 8 * 8), max(64, int(W * f) // 8 * 8)
        x = O.make_input("uniform", 1, sample_H, sample_W, seed=1234)
    for _ in range(args.warmup):
        O.forward(x, sd, run_dead_block=True, mssa=mssa)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.forward(x, sd, run_dead_block=True, mssa=mssa)
    dt = time.perf_counter() - t0
    mp = sample_H * sample_W / 1e6
    val = mp * args.steps / dt
    sample = f"{args.steps} forwards of 1x3x{sample_H}x{sample_W} (fp32, torch CPU, {torch.get_num_threads()} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (" [MSSA variant, net/CIDNet_MSSA.py]" if mssa else ""), "weights": wdesc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(sd, H, W, seconds=12.0, mssa=False):
    import torch
    from oracle import cidnet_oracle as O
    O.FAST_BILINEAR = True
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h, w = min(H, 400), min(W, 600)          # cfg1 frame: the reference's own CPU-runnable case
    x = O.make_input("uniform", 1, h, w, seed=1234)
    sdf = {k: v.float() for k, v in sd.items()}
    with torch.no_grad():
        O.forward(x, sdf, run_dead_block=True, mssa=mssa)
        n, t0 = 0, time.perf_counter()
        while True:
            O.forward(x, sdf, run_dead_block=True, mssa=mssa)
            n += 1
            dt = time.perf_counter() - t0
            if dt > seconds or n >= 20:
                break
    return {"value": h * w * n / dt / 1e6, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} forwards of 1x3x{h}x{w}, fp32 torch CPU (oracle port incl. dead I_LCA5), {dt:.1f} s"}


def gpu_eager_baseline(B, H, W, mssa=False, steps=10):
    """The reference's own GPU path as a bar (BASELINE.md 4.2 / SURVEY 8d): the oracle port of the reference forward --
    the identical ATen op sequence, incl. the dead I_LCA5 -- run eagerly with torch on this GPU, in PyTorch's default
    numeric mode (cuDNN convs in TF32) and in strict fp32.  Rank 0 only, outside every timed region of our arm."""
    import torch
    from oracle import cidnet_oracle as O
    O.FAST_BILINEAR = True
    dev = torch.device("cuda", torch.cuda.current_device())
    sd = {k: v.to(dev) for k, v in O.make_state_dict(0, False, mssa=mssa).items()}
    x = torch.rand(B, 3, H, W, device=dev)
    out = {"impl": "oracle port of the reference forward, torch eager on cuda (cuDNN / cuBLAS)", "steps": steps,
           "shape": [B, 3, H, W]}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for mode, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            for _ in range(3):
                O.forward(x, sd, run_dead_block=True, mssa=mssa)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.forward(x, sd, run_dead_block=True, mssa=mssa)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode + "_ms"] = ms
            out[mode + "_MPps"] = B * H * W / ms / 1e3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["MP/s"] = out["tf32_MPps"]
    return out


def measure_cfg5_sharded(model, dev, world, rank, steps, warmup, parity=True):
    """ONE 4K image, rows sharded over the ranks (hvi-cidnet_b200/dist.py RowShardedCIDNet -> cidnet_forward_sharded):
    conv halos and the partial Gram sums cross the GPUs over NCCL, kernels + exchanges replayed as one CUDA graph.
    Returns a dict (same on every rank): device-timed ms/step (max over ranks), MP/s, exchange counts / bytes, e2e
    (pinned strip -> H2D -> forward -> D2H of the owned rows), and -- computed inside this run -- the max-abs
    difference of the gathered sharded result to the UNSHARDED forward of the same image on rank 0."""
    import torch
    import torch.distributed as dist
    from hvi_cidnet_b200.dist import RowShardedCIDNet, strip_plan, strip_local_range
    B, H, W, desc = WORKLOADS["cfg5"]
    net = RowShardedCIDNet(model, halo=16, graph=True, transport=os.environ.get("CIDNET_SHARD_TRANSPORT"))   # default: peer memory
    sh = strip_plan(H, world, rank, 16)
    a, b = strip_local_range(sh)
    g = torch.Generator().manual_seed(1234)
    nimg = 3
    hfull = [torch.rand(1, 3, H, W, generator=g) for _ in range(nimg)]           # same images on every rank
    hloc = [t[:, :, a:b, :].contiguous().pin_memory() for t in hfull]
    xs = [t.to(dev) for t in hloc]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    nwarm, t0 = 0, time.perf_counter()
    while nwarm < max(3, warmup) or time.perf_counter() - t0 < 0.4:
        net.forward_strip(xs[nwarm % nimg], H)
        nwarm += 1
    barrier()
    sampler = ClockSampler(dev.index)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        y, _ = net.forward_strip(xs[i % nimg], H)
    e1.record()
    sampler.poll_until(e1)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if net.transport == "peer":
        # the schedule is the library's (the same as the callback transport's): count it with the host-only dry run
        import ctypes as C
        from hvi_cidnet_b200 import _lib
        L = _lib.lib()
        nb = L.cidnet_workspace_bytes(1, b - a, W)
        nh, na = C.c_int(), C.c_int()
        dummy = (C.c_char * 2048)()
        noop_h = _lib.HALO_FN(lambda u, r, n: 0)
        noop_a = _lib.ALLREDUCE_FN(lambda u, p, c: 0)
        base = (C.addressof(dummy) + 1023) & ~1023
        L.cidnet_forward_sharded_dry(W, C.byref(sh), C.c_void_p(base), C.c_int64(1 << 62), noop_h, noop_a, None, C.byref(nh), C.byref(na))
        halo_calls, ar_calls, sent = nh.value, na.value, None
    else:
        halo_calls = sum(1 for e in net.comm.log if e[0] == "halo")
        ar_calls = sum(1 for e in net.comm.log if e[0] == "allreduce")
        sent = net.comm.bytes_sent
    # end to end: pinned strip -> H2D -> sharded forward -> D2H of the owned rows
    hy = torch.empty(1, 3, sh.row_end - sh.row_begin, W).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(steps):
        y, _ = net.forward_strip(hloc[i % nimg].to(dev, non_blocking=True), H)
        hy.copy_(y[:, :, sh.row_begin - a:sh.row_end - a, :], non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    launches_per_step = model.num_launches()
    # in-run parity: gather the owned rows of image 0 on every rank, rank 0 compares with its own unsharded forward
    par = None
    if parity:
        y, _ = net.forward_strip(xs[0], H)
        own = y[:, :, sh.row_begin - a:sh.row_end - a, :].contiguous()
        plans = [strip_plan(H, world, r, 16) for r in range(world)]
        nmax = max(p.row_end - p.row_begin for p in plans)
        pad = own.new_zeros((1, 3, nmax, W))
        pad[:, :, :own.shape[2]] = own
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        pv = torch.zeros(2, device=dev, dtype=torch.float64)
        if rank == 0:
            full = torch.cat([bf[:, :, :p.row_end - p.row_begin] for bf, p in zip(bufs, plans)], dim=2)
            ref = model(hfull[0].to(dev))
            d = (full - ref).abs().amax(dim=1).flatten()
            top = torch.topk(d, 4).values
            # the reference's PHVIT black-pixel hole (tests/conftest.py::parity_error): report the pixels beyond 5e-4
            # and the max over the rest; a healthy run has 0 such pixels
            pv[0] = float(d[d <= 5e-4].max()) if bool((d <= 5e-4).any()) else float(top[0])
            pv[1] = float((d > 5e-4).sum())
            del full, ref
        dist.broadcast(pv, 0)
        par = {"parity_vs_unsharded_maxabs": float(pv[0]), "pixels_beyond_5e-4": int(pv[1])}
    mp_step = H * W / 1e6
    how = ("peer-memory transport: halo rows / attention statistics read from the neighbours' HBM over NVLink by the library's "
           "own kernels, one CUDA graph per strip" if net.transport == "peer" else
           f"NCCL send/recv + all-reduce from host callbacks, CUDA-graph replay of kernels + NCCL exchanges ({net.replays} replays)")
    res = {"workload": f"cfg5: CIDNet 1x3x{H}x{W}, rows sharded over {world} GPUs (halo 16 rows), {how}",
           "transport": net.transport, "peer_error": net.peer_error() if net.transport == "peer" else None,
           "ms_per_step": ms_total / steps, "MP/s": mp_step * steps / (ms_total / 1e3), "steps": steps, "warmup": nwarm,
           "halo_calls": halo_calls, "allreduce_calls": ar_calls, "bytes_sent": sent,
           "e2e_ms_per_step": ms_e2e / steps, "e2e_MP/s": mp_step * steps / (ms_e2e / 1e3),
           "h2d_bytes_per_step": hloc[0].numel() * 4, "d2h_bytes_per_step": hy.numel() * 4,
           "local_rows_rank0": strip_local_range(strip_plan(H, world, 0, 16))[1], "graph_replays": net.replays,
           "graph_error": getattr(net, "graph_error", None), "launches_per_step": launches_per_step, "clocks": clocks}
    if par:
        res.update(par)
    net.close()
    torch.cuda.synchronize()
    del net, xs
    torch.cuda.empty_cache()
    return res


def measure_cfg4_batch(model, dev, world, rank, steps, warmup):
    """64 images of 400x600 split contiguously over the ranks (strong scaling: fixed total work), no collective on the
    data path.  In-run parity: the first image of the rank's share, run alone, against its slot in the batched result
    (max over ranks)."""
    import torch
    import torch.distributed as dist
    B, H, W, desc = WORKLOADS["cfg4"]
    from hvi_cidnet_b200.dist import shard_range
    start, count = shard_range(B, world, rank)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    ring = 3
    xs = [torch.rand(count, 3, H, W, device=dev, generator=g) for _ in range(ring)]
    for i in range(max(3, min(warmup, 5))):
        model(xs[i % ring])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        y = model(xs[i % ring])
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    yb = model(xs[0])
    y1 = model(xs[0][0:1].contiguous())
    d = (yb[0:1] - y1).abs().amax(dim=1).flatten()
    t = torch.tensor([ms, float(d[d <= 5e-4].max()) if bool((d <= 5e-4).any()) else float(d.max()), float((d > 5e-4).sum())],
                     device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    del xs, y, yb, y1
    model._workspaces = {}
    torch.cuda.empty_cache()
    return {"workload": f"cfg4: CIDNet {B}x3x{H}x{W} batch split over {world} GPUs ({count} images on rank {rank})",
            "ms_per_step": ms / steps, "MP/s": B * H * W / 1e6 * steps / (ms / 1e3), "steps": steps, "scaling": "strong",
            "parity_batch_vs_single_maxabs": float(t[1]), "pixels_beyond_5e-4": int(t[2])}


def run_extras_under_watchdog(jobs, deadline, rank, line):
    """Run the extra workloads `jobs` = [(key, callable), ...] and return {key: result | {"error": ...}}.  If they are not
    done after `deadline` seconds (a rank that never arrives at a collective, a stalled exchange), rank 0 prints the ONE
    JSON line `line` with the results so far and the reason, and EVERY rank's process ends (os._exit(0)): the headline is
    never lost to an extra.  `line` is None on the other ranks."""
    extras = {}
    finished = threading.Event()
    print_once = threading.Lock()

    def watchdog():
        if finished.wait(deadline) or not print_once.acquire(blocking=False):
            return
        if rank == 0 and line is not None:
            extras.setdefault("error", f"extra workloads did not finish within {deadline:.0f} s (CIDNET_EXTRAS_TIMEOUT); "
                                       "the headline of this line is complete")
            line["extra_workloads"] = extras
            print(json.dumps(line), flush=True)
        os._exit(0)
    threading.Thread(target=watchdog, daemon=True).start()
    t0 = time.perf_counter()
    for key, fn in jobs:
        try:
            extras[key] = fn()
        except Exception as e:               # never lose the headline line to an extra
            extras[key] = {"error": repr(e)[:400]}
        print(f"[bench rank {rank}] extra {key} done after {time.perf_counter() - t0:.1f} s", file=sys.stderr, flush=True)
    finished.set()
    if not print_once.acquire(blocking=False):       # the watchdog is printing / has printed: it also ends the process
        time.sleep(30)
        os._exit(0)
    return extras


def run_ours(args):
    import torch
    import torch.distributed as dist
    if args.variant == "mssa":
        from hvi_cidnet_b200.net.CIDNet_MSSA import CIDNet
    else:
        from hvi_cidnet_b200.net.CIDNet import CIDNet

    torch.set_grad_enabled(False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CIDNet hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    extras_wanted = args.workload is None and world > 1 and not args.no_extras
    if args.workload is None:
        args.workload = "cfg2"
    B, H, W, desc = WORKLOADS[args.workload]
    if args.workload == "cfg4":
        B = max(1, B // world)               # batch-sharded: fixed total, per-rank share
    sd, wdesc = make_weights(mssa=args.variant == "mssa")
    model = CIDNet().to(dev).eval()
    model.load_state_dict(sd, strict=True)
    if args.workload == "cfg5" and world > 1:
        return run_cfg5_sharded(args, model, sd, wdesc, dev, world, rank, peaks)

    # inputs: a ring of distinct images whose total size exceeds L2, so no step finds its input
    # (or the previous step's intermediates, which are rewritten every step) in cache
    img_bytes = B * 3 * H * W * 4
    ring = max(2, min(64, -(-2 * L2_BYTES // img_bytes)))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.rand(B, 3, H, W, device=dev, generator=g) for _ in range(ring)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ---------------------------------------------------
    # warm-up: at least W steps AND at least ~0.4 s of back-to-back work, so the SM clocks have
    # ramped from idle before the timed region (the clocks line below records what was seen)
    nwarm, t_w0 = 0, time.perf_counter()
    while nwarm < max(3, args.warmup) or time.perf_counter() - t_w0 < 0.4:
        model(xs[nwarm % ring])
        nwarm += 1
        if nwarm % 8 == 0:
            torch.cuda.synchronize()
    # the timed region: EXACTLY K steps between two events, barrier + synchronize on both sides; repeated
    # three times back to back and the MEDIAN pass is reported (all pass times are kept in the JSON line):
    # single passes were occasionally ~2x slow with no clock change when an NVML query stalled the device
    passes = []
    for _ in range(3):
        barrier()
        sampler = ClockSampler(local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            y = model(xs[i % ring])
        e1.record()
        sampler.poll_until(e1)          # clocks / throttle reasons while the timed steps execute
        barrier()
        passes.append((e0.elapsed_time(e1), sampler.stop()))
    ms_total, clocks = sorted(passes, key=lambda p: p[0])[1]
    clocks["passes_ms_per_step"] = [round(p[0] / args.steps, 4) for p in passes]
    launches = model.num_launches() * args.steps

    # ---- per-kernel events, same steps (roofline) -----------------------------------------
    # The profiled pass runs eagerly with the SAME two-stream split as the replayed graph (the I and HV launches of a pair
    # overlap on half the SMs each), so a kernel's time is the one it has in the schedule `value` times.  Launches that
    # overlap in time are merged into one group: its wall time is max(end) - min(start), its bytes / flops the sum.
    model.set_profiling(True)
    agg = {}
    for i in range(args.steps):
        model(xs[i % ring])
        recs = model.read_profile(spans=True)
        groups = []
        for name, ms, by, fl, t0, t1 in recs:
            g = None
            for cand in groups[-4:]:              # the sibling launch of a pair is at most a few records back
                if cand["name"] == name and t0 < cand["t1"] - 0.25 * min(ms, cand["t1"] - cand["t0"]):
                    g = cand
            if g is not None:
                g["t1"] = max(g["t1"], t1); g["t0"] = min(g["t0"], t0); g["by"] += by; g["fl"] += fl; g["n"] += 1
            else:
                groups.append({"name": name, "t0": t0, "t1": t1, "by": by, "fl": fl, "n": 1})
        for g in groups:
            a = agg.setdefault(g["name"], [0.0, 0.0, 0.0, 0, 0])
            a[0] += g["t1"] - g["t0"]; a[1] += g["by"]; a[2] += g["fl"]; a[3] += g["n"]; a[4] += 1
    model.set_profiling(False)

    # ---- end to end through the public API with HOST buffers --------------------------------
    hx = [torch.rand(B, 3, H, W).pin_memory() for _ in range(min(ring, 4))]
    hy = torch.empty(B, 3, H, W).pin_memory()
    for i in range(3):
        hy.copy_(model(hx[i % len(hx)].to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        xin = hx[i % len(hx)].to(dev, non_blocking=True)
        hy.copy_(model(xin), non_blocking=True)
    f1.record()
    barrier()
    ms_e2e_sync = f0.elapsed_time(f1)
    # the same through the streamed driver (hvi-cidnet_b200/stream.py): H2D / forward / D2H of consecutive
    # steps overlap on three streams; every step still uploads its input and downloads its result
    from hvi_cidnet_b200.stream import StreamedCIDNet
    drv = StreamedCIDNet(model, depth=3)
    for _ in drv.run(hx[i % len(hx)] for i in range(3)):
        pass
    barrier()
    f0.record()
    nres = 0
    for res in drv.run(hx[i % len(hx)] for i in range(args.steps)):
        nres += 1
    f1.record()
    barrier()
    assert nres == args.steps
    ms_e2e = f0.elapsed_time(f1)
    # the 8-bit path: uint8 HWC in pinned host memory -> H2D -> forward with the conversions fused into its first / last
    # kernel -> D2H of the uint8 result (6 B/px over PCIe instead of 24)
    h8 = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(min(ring, 4))]
    for _ in drv.run_u8(h8[i % len(h8)] for i in range(3)):
        pass
    barrier()
    f0.record()
    nres = 0
    for res in drv.run_u8(h8[i % len(h8)] for i in range(args.steps)):
        nres += 1
    f1.record()
    barrier()
    ms_e2e_u8 = f0.elapsed_time(f1)
    del h8

    del xs, hx, hy, drv
    model._workspaces = {}
    torch.cuda.empty_cache()
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_e2e_sync, ms_e2e_u8], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_e2e_sync, ms_e2e_u8 = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    mp_step_all = B * H * W * world / 1e6
    value = mp_step_all * args.steps / (ms_total / 1e3)
    e2e_value = mp_step_all * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        kern = []
        tot = sum(a[0] for a in agg.values())
        for name, (ms, by, fl, n, ng) in agg.items():
            kern.append({"name": name, "launches_per_step": n // args.steps, "concurrent_groups_per_step": ng // args.steps,
                         "ms_per_step": ms / args.steps,
                         "share": ms / tot if tot else 0.0,
                         "GBps": by / ms / 1e6 if ms else 0.0, "TFLOPs": fl / ms / 1e9 if ms else 0.0})
        kern.sort(key=lambda k: -k["ms_per_step"])
        top = kern[0]
        roof = {"kernel": top["name"], "bound": "hbm", "achieved": top["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": top["GBps"] / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "share_of_step": top["share"], "TFLOPs": top["TFLOPs"],
                "cuda_core_fp32_peak_TFLOPs": 148 * 128 * 2 * 1.965e-3,    # 148 SMs x 128 FMA lanes x 2 x 1.965 GHz (HFMA2 issues at the same FMA rate)
                "note": "achieved = algorithmic bytes (DESIGN.md) / CUDA-event wall time of that kernel inside the forward, in the "
                        "two-stream schedule of the replayed graph (the overlapping I / HV launches of a pair count as one group)"}
        # DRAM traffic of the same kernel from the committed ncu capture of this workload (per launch)
        tpath = os.path.join(ROOT, "profiles", f"r02_traffic_{args.workload}.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", f"r01_traffic_{args.workload}.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath))["per_launch"].get(top["name"])
            if tr:
                roof["traffic"] = tr["dram_bytes"]
                roof["traffic_source"] = os.path.relpath(tpath, ROOT)
                roof["algorithmic_bytes_per_launch"] = tr["algorithmic_bytes"]
        cpu = cpu_baseline_sample(sd, H, W, mssa=args.variant == "mssa")
        try:
            eager = gpu_eager_baseline(min(B, 8), H, W, mssa=args.variant == "mssa")
        except Exception as e:
            eager = {"error": repr(e)[:300]}
        act = "fp16" if __import__("hvi_cidnet_b200._lib", fromlist=["lib"]).lib().cidnet_act_dtype() == 0 else "bf16"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": nwarm,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload != "cfg4" else "strong",
                "vs_baseline": None, "dtype": act, "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}" + (" [MSSA variant, net/CIDNet_MSSA.py]" if args.variant == "mssa" else ""), "per_rank_batch": B, "H": H, "W": W, "weights": wdesc,
                           "l2": f"inputs rotate over a ring of {ring} distinct images ({ring * img_bytes >> 20} MiB > L2); "
                                 "all intermediates are rewritten every step",
                           "accumulate": "fp32", "parallelism": f"dp{world} (independent images, no collective)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": img_bytes, "d2h_bytes_per_step": img_bytes,
                        "ms_per_step": ms_e2e / args.steps, "api": "StreamedCIDNet(model).run(pinned host batches) -> pinned host results",
                        "sync_loop_value": mp_step_all * args.steps / (ms_e2e_sync / 1e3),
                        "sync_loop_note": "reference-style loop: x.cuda() -> model(x) -> .cpu() per step on one stream",
                        "u8_value": mp_step_all * args.steps / (ms_e2e_u8 / 1e3), "u8_bytes_per_step_each_way": B * H * W * 3,
                        "u8_api": "StreamedCIDNet(model).run_u8(pinned uint8 HWC batches): ToTensor / pad / gamma / clamp / crop / "
                                  "quantise fused into the stem and head kernels (cidnet_forward_u8)"},
                "gpu_launches": launches, "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
                "gpu_eager_baseline": eager, "clocks": clocks}
    # ---- the two multi-GPU configs of BASELINE.json, same process group (N > 1, default workload only) -------------
    # They run AFTER the headline line is complete and under a watchdog: whatever happens to an extra (an exception, a
    # rank that never arrives), rank 0 prints the ONE JSON line -- with what the extras produced so far and the reason --
    # within CIDNET_EXTRAS_TIMEOUT seconds, and every rank leaves.
    if extras_wanted:
        extras = run_extras_under_watchdog(
            [("cfg5", lambda: measure_cfg5_sharded(model, dev, world, rank, steps=min(args.steps, 10), warmup=3)),
             ("cfg4", lambda: measure_cfg4_batch(model, dev, world, rank, steps=min(args.steps, 5), warmup=3))],
            float(os.environ.get("CIDNET_EXTRAS_TIMEOUT", "180")), rank, line if rank == 0 else None)
    if rank == 0:
        if extras_wanted:
            line["extra_workloads"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg5_sharded(args, model, sd, wdesc, dev, world, rank, peaks):
    """--workload cfg5 --gpus N>1 (BASELINE.json configs[4]) as the headline line: see measure_cfg5_sharded."""
    import torch.distributed as dist
    B, H, W, desc = WORKLOADS["cfg5"]
    r = measure_cfg5_sharded(model, dev, world, rank, args.steps, args.warmup)
    if rank == 0:
        act = "fp16" if __import__("hvi_cidnet_b200._lib", fromlist=["lib"]).lib().cidnet_act_dtype() == 0 else "bf16"
        line = {"metric": METRIC, "value": r["MP/s"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": act, "data": "synthetic",
                "config": {"workload": r["workload"], "H": H, "W": W, "weights": wdesc, "local_rows_rank0": r["local_rows_rank0"],
                           "l2": "3 distinct images in rotation; all intermediates are rewritten every step",
                           "parallelism": f"spatial row strips x{world}: {r['halo_calls']} halo exchanges + {r['allreduce_calls']} Gram all-reduces "
                                          f"per forward over NCCL, {r['bytes_sent']} halo bytes sent per rank per forward"},
                "e2e": {"value": r["e2e_MP/s"], "unit": UNIT, "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": r["d2h_bytes_per_step"], "ms_per_step": r["e2e_ms_per_step"]},
                "gpu_launches": r["launches_per_step"] * args.steps, "roofline": None, "cpu_baseline": None,
                "parity_vs_unsharded_maxabs": r.get("parity_vs_unsharded_maxabs"), "clocks": r["clocks"]}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def run_hvi(args):
    """--workload cfg3: the standalone RGB<->HVI transform (BASELINE.json configs[2]).  A step = HVIT then
    PHVIT over the whole batch (48 B/pixel algorithmic: 24 per direction, fp32 planar)."""
    import torch
    import torch.distributed as dist
    from hvi_cidnet_b200.net.HVI_transform import RGB_HVI
    from oracle import cidnet_oracle as O
    torch.set_grad_enabled(False)
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B, H, W, desc = WORKLOADS["cfg3"]
    t = RGB_HVI().to(dev)
    x = torch.rand(B, 3, H, W, device=dev)          # 796 MB >> L2: every step streams from HBM
    nbytes = x.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nwarm, t0 = 0, time.perf_counter()
    while nwarm < max(3, args.warmup) or time.perf_counter() - t0 < 0.4:
        y = t.PHVIT(t.HVIT(x)); nwarm += 1
        torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        hvi = t.HVIT(x); ev[2 * i + 1].record()
        y = t.PHVIT(hvi); ev[2 * i + 2].record()
    sampler.poll_until(ev[-1])
    barrier()
    clocks = sampler.stop()
    ms_total = ev[0].elapsed_time(ev[-1])
    ms_h = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
    ms_p = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)) / args.steps
    hx = torch.rand(B, 3, H, W).pin_memory(); hy = torch.empty(B, 3, H, W).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(max(2, args.steps // 4)):
        hy.copy_(t.PHVIT(t.HVIT(hx.to(dev, non_blocking=True))), non_blocking=True)
    f1.record(); barrier()
    ms_e2e = f0.elapsed_time(f1) / max(2, args.steps // 4)
    # backward of the transform (train.py:61-62): one launch each, 36 B/pixel (input, upstream gradient, result)
    from hvi_cidnet_b200 import _lib
    L = _lib.lib()
    go = torch.randn(B, 3, H, W, device=dev); gi = torch.empty_like(x)
    gk = torch.empty(1, device=dev); scratch = torch.empty(int(L.cidnet_hvi_backward_scratch_bytes()) // 4, device=dev)
    kd = t.density_k.detach().float().contiguous()
    sp = _lib.stream_ptr(dev)
    bw = {}
    for name, call in (("hvit_backward", lambda: L.cidnet_hvit_backward(x.data_ptr(), go.data_ptr(), gi.data_ptr(), gk.data_ptr(),
                                                                         scratch.data_ptr(), B, H, W, 0.0, kd.data_ptr(), sp)),
                       ("phvit_backward", lambda: L.cidnet_phvit_backward(hvi.data_ptr(), go.data_ptr(), gi.data_ptr(), B, H, W, 0.0,
                                                                           kd.data_ptr(), 0, 1.3, 0, 1.0, sp))):
        for _ in range(3):
            _lib.check(call())
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); b0.record()
        for _ in range(args.steps):
            _lib.check(call())
        b1.record(); barrier()
        ms_b = b0.elapsed_time(b1) / args.steps
        bw[name] = {"ms": ms_b, "GBps": nbytes * 3 / ms_b / 1e6, "frac_of_hbm_peak": nbytes * 3 / ms_b / 1e6 / peaks["hbm_gbs"],
                    "algorithmic_bytes_per_px": 36, "launches": 2 if name == "hvit_backward" else 1}
    # the same backward as torch autograd derives it from the reference's formulation (oracle port, eager on this GPU)
    with torch.enable_grad():
        xe = x[:2].clone().requires_grad_(True)
        ke = torch.full([1], 0.2, device=dev, requires_grad=True)
        r_, g_, b_ = xe[:, 0], xe[:, 1], xe[:, 2]
        for it in range(3):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            val = xe.max(1)[0]; vmin = xe.min(1)[0]; d = val - vmin + 1e-8
            hue = torch.where(b_ == val, 4.0 + (r_ - g_) / d, torch.zeros_like(val))
            hue = torch.where(g_ == val, 2.0 + (b_ - r_) / d, hue)
            hue = torch.where(r_ == val, torch.remainder((g_ - b_) / d, 6), hue)
            hue = torch.where(vmin == val, torch.zeros_like(hue), hue) / 6.0
            sat = torch.where(val == 0, torch.zeros_like(val), (val - vmin) / (val + 1e-8))
            cs = ((val * 0.5 * math.pi).sin() + 1e-8).pow(ke)
            out = torch.stack([cs * sat * (2.0 * math.pi * hue).cos(), cs * sat * (2.0 * math.pi * hue).sin(), val], dim=1)
            e1.record()
            out.backward(go[:2])
            e2.record(); torch.cuda.synchronize()
            xe.grad = None; ke.grad = None
        bw["eager_autograd_hvit"] = {"batch": 2, "forward_ms": e0.elapsed_time(e1), "backward_ms": e1.elapsed_time(e2),
                                     "backward_MPps": 2 * H * W / e1.elapsed_time(e2) / 1e3,
                                     "ours_backward_MPps": B * H * W / bw["hvit_backward"]["ms"] / 1e3}
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_total, ms_e2e = float(tt[0]), float(tt[1])
    if rank == 0:
        mp = B * H * W * world / 1e6
        xs = O.make_input("uniform", 2, H, W, seed=1)
        t1 = time.perf_counter(); n = 0
        while time.perf_counter() - t1 < 8.0:
            O.phvit(O.hvit(xs, 0.2), 0.2); n += 1
        cpu_v = 2 * H * W * n / (time.perf_counter() - t1) / 1e6
        gb_h, gb_p = nbytes * 2 / ms_h / 1e6, nbytes * 2 / ms_p / 1e6
        worst = min(gb_h, gb_p)
        line = {"metric": "RGB<->HVI round trip megapixels/s", "value": mp * args.steps / (ms_total / 1e3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": nwarm, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg3: " + desc, "per_rank_batch": B, "l2": "796 MB per tensor, far larger than L2"},
                "e2e": {"value": mp / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
                "gpu_launches": 2 * args.steps,
                "roofline": {"kernel": "hvi_vec4_kernel<PHVIT>" if gb_p < gb_h else "hvi_vec4_kernel<HVIT>", "bound": "hbm",
                             "achieved": worst, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": worst / peaks["hbm_gbs"],
                             "traffic": None, "peak_source": peaks["source"],
                             "hvit_GBps": gb_h, "phvit_GBps": gb_p, "algorithmic_bytes_per_px": 24},
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{n} round trips of 2x3x{H}x{W}, torch CPU fp32 (oracle port)"},
                "backward": bw,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: cfg2 (BASELINE.json configs[1]); with --gpus N > 1 and no --workload the line also carries "
                         "`extra_workloads` (cfg5 row-sharded, cfg4 batch-split)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_workloads of a default multi-GPU run")
    ap.add_argument("--variant", default="base", choices=["base", "mssa"],
                    help="base = net/CIDNet.py (BASELINE.json's model); mssa = the fork's net/CIDNet_MSSA.py")
    args = ap.parse_args()
    if args.impl == "reference":
        args.workload = args.workload or "cfg2"
        run_reference(args)
    elif args.workload == "cfg3":
        run_hvi(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- sparse-conv voxels/sec forward+backward over a GridBatch (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c2f32|c2x128|c3|c4|c5|c5f32]

One "step" = one pass of the hot path over one batch: forward + dgrad + wgrad of a 3^3 64->64 bf16
SparseConv3d-shaped plan (BASELINE.json configs[1]: 8 indoor grids x ~200 k voxels, same-topology target),
kernel map prebuilt ("topology amortized", as the reference's own benches do).  For N > 1 (launched under
torch.distributed.run) every rank owns whole grids (its own batch of 8: weak scaling; --config c4 partitions ONE 32-grid
batch by grid: strong scaling) and the step contains the NCCL all-reduce of grad_weights, the path's only exchange, issued
asynchronously behind wgrad so that it overlaps dgrad.  --config c3 runs the sparse UNet block-stack training step.

Prints ONE JSON line on rank 0; see the task contract for the keys.  `value` is timed with inputs
resident in HBM; `e2e` goes through the same C-ABI-backed calls with HOST (pinned) buffers, H2D and D2H
copies inside the timed region.  `--impl reference` times the CPU restatement of the reference's own
CPU path (oracle/, torch::mm semantics, all host threads) on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for p in (str(REPO), str(REPO / "fvdb-core_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    # name: (generator, grids, target voxels per grid, kernel, cin, cout, dtype)
    "c1": dict(gen="sphere_shell", grids=1, voxels=100_000, kernel=3, cin=32, cout=32, dtype="f32", desc="C1 single grid ~100k voxels, 3^3 32->32 fp32"),
    "c2": dict(gen="indoor_room", grids=8, voxels=200_000, kernel=3, cin=64, cout=64, dtype="bf16", desc="C2 ScanNet-shaped 8 grids x ~200k voxels, 3^3 64->64 bf16 fwd+bwd"),
    "c3": dict(gen="indoor_room", grids=16, voxels=150_000, kernel=3, cin=32, cout=32, dtype="bf16", desc="C3 sparse UNet block stack (3^3 convs, 2^3 s2 down, transposed up, 32..256 ch) on 16 indoor grids, training step"),
    "c2f32": dict(gen="indoor_room", grids=8, voxels=200_000, kernel=3, cin=64, cout=64, dtype="f32", desc="C2-shaped 8 grids x ~200k voxels, 3^3 64->64 fp32 fwd+bwd (three-way bf16 split on the tensor pipe)"),
    "c2x128": dict(gen="indoor_room", grids=8, voxels=200_000, kernel=3, cin=128, cout=128, dtype="bf16", desc="C2-shaped 8 grids x ~200k voxels, 3^3 128->128 bf16 fwd+bwd"),
    "c4": dict(gen="lidar_sweep", grids=32, voxels=1_000_000, kernel=3, cin=128, cout=128, dtype="bf16", partition="by_grid",
               desc="C4 KITTI-shaped 32 grids x ~1M voxels, 3^3 128->128 bf16, GridBatch partitioned by grid"),
    "c5": dict(gen="random_occupancy", grids=8, voxels=4_979_000, kernel=5, cin=16, cout=16, dtype="bf16", desc="C5 8 grids x ~5M voxels, 5^3 16->16"),
    "c5f32": dict(gen="random_occupancy", grids=8, voxels=4_979_000, kernel=5, cin=16, cout=16, dtype="f32", desc="C5 8 grids x ~5M voxels, 5^3 16->16 fp32 (three-way bf16 split on the tensor pipe)"),
}
DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16, "f64": torch.float64}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of the same command
# (profiles/r01_ncu_full_final_summary.txt); only known for the default workload.
NCU_TRAFFIC_BYTES = {"c2": {"fwd": 297.7e6 + 170.9e6, "dgrad": 297.7e6 + 170.9e6, "wgrad": 947.3e6 + 15.0e6}}


def load_peaks() -> dict:
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def make_coords(cfg: dict, rank: int, device, world: int = 1) -> list[torch.Tensor]:
    """This rank's grids.  Weak scaling (default): every rank builds its own cfg["grids"] grids.  cfg["partition"] ==
    "by_grid" (C4): the batch of cfg["grids"] grids is ONE job partitioned by grid -- every rank generates the whole
    batch (same seeds), bin-packs it by voxel count (LPT) and keeps its share, so no data crosses ranks."""
    from fvdb.utils import synthetic

    gen = getattr(synthetic, cfg["gen"])
    if cfg.get("partition") == "by_grid" and world > 1:
        from fvdb.distributed import partition_grids_lpt

        every = make_coords({**cfg, "partition": None}, 0, device, 1)
        mine = partition_grids_lpt([int(c.shape[0]) for c in every], world)[rank]
        return [every[g] for g in mine]
    out = []
    for g in range(cfg["grids"]):
        seed = rank * 1000 + g
        if cfg["gen"] == "random_occupancy":
            out.append(gen(seed=42 + seed, device=device))
        elif cfg["gen"] == "lidar_sweep":  # ray casting vectorised on the device (scene parameters from a CPU generator)
            out.append(gen(target=cfg["voxels"], seed=seed, device=device))
        else:
            out.append(gen(target=cfg["voxels"], seed=seed, device="cpu").to(device))
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([f.strip() for f in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(P, n_in, n_out, cin, cout, k3, s):
    """gathered bytes per pass."""
    return {
        "fwd": P * cin * s + n_out * cout * s + 4 * P + k3 * cin * cout * s,
        "dgrad": P * cout * s + n_in * cin * s + 4 * P + k3 * cin * cout * s,
        "wgrad": P * (cin + cout) * s + 8 * P + 4 * k3 * cin * cout,
    }


# ------------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference's CPU path on a bounded sample
# ------------------------------------------------------------------------------------------------------


def cpu_reference_step_time(cfg: dict, sample_grids: int, repeats: int, warmup: int):
    """Seconds per fwd+bwd of the oracle (per-tap index_select -> mm -> index_add_, fp32, all host threads)
    on `sample_grids` grids of the workload.  Returns (seconds per step, voxels per step, pairs, threads)."""
    import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    coords = make_coords({**cfg, "grids": sample_grids}, 0, "cpu")
    ijk = torch.cat(coords).numpy().astype(np.int64)
    bidx = np.concatenate([np.full(len(c), i, dtype=np.int64) for i, c in enumerate(coords)])
    order = oracle.index_grid_row_order(bidx, ijk)
    ijk, bidx = ijk[order], bidx[order]
    k = cfg["kernel"]
    topo = oracle.build_topology(ijk, bidx, ijk, bidx, k, 1)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((len(ijk), cfg["cin"]), generator=gen)
    w = (torch.rand((cfg["cout"], cfg["cin"], k, k, k), generator=gen) * 2 - 1) / (cfg["cin"] * k**3) ** 0.5
    dy = torch.randn((len(ijk), cfg["cout"]), generator=gen)
    times, begin = [], time.perf_counter()
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        oracle.gs_conv(x, w, topo)
        oracle.gs_conv_backward(dy, x, w, topo)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        if times and time.perf_counter() - begin > 150.0:  # keep the whole run within a few minutes
            break
    return float(np.median(times)), len(ijk), topo.total_pairs, torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_grids = 1
    sec, voxels, pairs, threads = cpu_reference_step_time(cfg, sample_grids, max(1, args.steps), max(0, min(args.warmup, 2)))
    value = voxels / sec
    line = {
        "impl": "reference", "metric": "sparse-conv voxels/sec fwd+bwd", "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": cfg["desc"], "sample": f"{sample_grids} of {cfg['grids']} grids per step ({voxels} voxels, {pairs} pairs)"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_grids} grid ({voxels} voxels) fwd+bwd fp32 per step, oracle port of GatherScatterDefault CPU path"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------


def has_fixed_topology(plan) -> bool:
    return bool(plan.has_fixed_topology)


def run_ours(args, cfg):
    import torch.distributed as dist

    import fvdb
    from fvdb import _fvdb_cpp as cpp
    from fvdb._lib import launch_count

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dtype = DTYPES[cfg["dtype"]]
    k, cin, cout = cfg["kernel"], cfg["cin"], cfg["cout"]
    coords = make_coords(cfg, rank, dev, world)
    strong = cfg.get("partition") == "by_grid"
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid)
    topo = plan._backend.topology
    topo._dgrad_plan()  # the input-stationary map for dgrad is part of the (amortised) plan (a symmetric map serves as it is)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    n, P, k3 = grid.total_voxels, topo.total_pairs, topo.kernel_volume

    gen = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn((n, cin), generator=gen).to(dtype).pin_memory()
    dy_host = torch.randn((n, cout), generator=gen).to(dtype).pin_memory()
    w = ((torch.rand((cout, cin, k, k, k), generator=gen) * 2 - 1) / (cin * k**3) ** 0.5).to(dtype).to(dev)
    x, dy = x_host.to(dev), dy_host.to(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    phase_ms = {"fwd": [], "dgrad+wgrad": []}

    def step(record: bool):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        y = cpp.gs_conv(x, w, topo)
        e1.record()
        pending = []
        # the step's only exchange: the all-reduce of grad_weights starts as soon as wgrad is enqueued and overlaps dgrad
        gx, gw = cpp.gs_conv_backward(dy, x, w, topo, on_grad_weights=(lambda g: pending.append(dist.all_reduce(g, async_op=True))) if world > 1 else None)
        for work in pending:
            work.wait()
        e2.record()
        if record:
            phase_ms["fwd"].append((e0, e1))
            phase_ms["dgrad+wgrad"].append((e1, e2))
        return y, gx, gw

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    launches0 = launch_count()
    with ClockSampler(local_rank) as clocks:
        start, stop = ev(), ev()
        start.record()
        for _ in range(args.steps):
            step(True)
        stop.record()
        barrier()
    launches = launch_count() - launches0
    ms = start.elapsed_time(stop) / args.steps
    fwd_ms = float(np.mean([a.elapsed_time(b) for a, b in phase_ms["fwd"]]))
    bwd_ms = float(np.mean([a.elapsed_time(b) for a, b in phase_ms["dgrad+wgrad"]]))

    # per-kernel durations of the three hot kernels, each timed alone on the launching stream
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    code = cpp._DTYPE_CODE[dtype]
    w_fwd = cpp._pack_weights(w, dtype, 0)
    in_map, in_mask, mirror = topo._dgrad_plan()
    w_bwd = cpp._pack_weights(w, dtype, 1, flip_taps=mirror)
    out_map = topo._out_map()
    kern_ms = {
        "fwd": timed(lambda: cpp._run_conv(x, w_fwd, out_map, n, n, cin, cout, k3, None, topo._out_mask()), max(3, args.steps)),
        "dgrad": timed(lambda: cpp._run_conv(dy, w_bwd, in_map, n, n, cout, cin, k3, None, in_mask), max(3, args.steps)),
    }
    kern_ms["wgrad"] = max(bwd_ms - kern_ms["dgrad"], 1e-6)  # (with N > 1 this includes the un-overlapped tail of the all-reduce)
    peaks = load_peaks()
    s = x.element_size()
    abytes = algorithmic_bytes(P, n, n, cin, cout, k3, s)
    flops = 2.0 * P * cin * cout
    dominant = max(kern_ms, key=kern_ms.get)
    per_kernel = {}
    for name in kern_ms:
        t = kern_ms[name] * 1e-3
        hbm_t, tensor_t = abytes[name] / (peaks["hbm_gbs"] * 1e9), flops / (peaks["tflops"] * 1e12)
        bound = "hbm" if hbm_t >= tensor_t else "tensor"
        achieved = abytes[name] / t / 1e9 if bound == "hbm" else flops / t / 1e12
        peak = peaks["hbm_gbs"] if bound == "hbm" else peaks["tflops"]
        per_kernel[name] = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": achieved / peak,
                            "ms": kern_ms[name], "algorithmic_bytes": abytes[name], "flops": flops}
    roof_time = sum(max(abytes[nm] / (peaks["hbm_gbs"] * 1e9), flops / (peaks["tflops"] * 1e12)) for nm in abytes)

    # end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    y_host = torch.empty((n, cout), dtype=dtype).pin_memory()
    gx_host = torch.empty((n, cin), dtype=dtype).pin_memory()
    gw_host = torch.empty(tuple(w.shape), dtype=dtype).pin_memory()

    # three streams: host->device copies, kernels, device->host copies (PCIe is full duplex, so the read-back of
    # y overlaps the upload of grad_out, and kernels overlap both)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def e2e_step():
        main = torch.cuda.current_stream(dev)
        s_in.wait_stream(main)
        with torch.cuda.stream(s_in):
            xd = x_host.to(dev, non_blocking=True)
            x_ready = torch.cuda.Event()
            x_ready.record(s_in)
            dyd = dy_host.to(dev, non_blocking=True)
            dy_ready = torch.cuda.Event()
            dy_ready.record(s_in)
        main.wait_event(x_ready)
        y = cpp.gs_conv(xd, w, topo)
        y_done = torch.cuda.Event()
        y_done.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(y_done)
            y_host.copy_(y, non_blocking=True)
        main.wait_event(dy_ready)
        gx, gw = cpp.gs_conv_backward(dyd, xd, w, topo)
        if world > 1:
            dist.all_reduce(gw)
        g_done = torch.cuda.Event()
        g_done.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(g_done)
            gx_host.copy_(gx, non_blocking=True)
            gw_host.copy_(gw, non_blocking=True)
        main.wait_stream(s_out)  # the step ends when its results are in host memory
        for t in (xd, dyd):
            t.record_stream(main)  # allocated on the copy-in stream, consumed by kernels on the main stream
        for t in (y, gx, gw):
            t.record_stream(s_out)  # allocated on the main stream, read by the copy-out stream

    e2e_mode = "3-stream overlap, whole batch"
    if cpp.lib.fvc_conv_scratch_bytes(n, n, cin, cout, k3, code) > 0 and dtype != torch.float32 and has_fixed_topology(plan):  # tensor-core path available
        from fvdb.streaming import HostPipelinedConv

        pipe = HostPipelinedConv(plan, num_chunks=args.e2e_chunks)
        e2e_mode = f"3-stream overlap, {len(pipe.bounds)} row chunks (fvdb.streaming.HostPipelinedConv)"

        def e2e_step():  # noqa: F811
            pipe.forward_backward(x_host, dy_host, w, y_host, gx_host, gw_host, reduce_fn=(lambda g: dist.all_reduce(g)) if world > 1 else None)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    a, b = ev(), ev()
    a.record()
    for _ in range(args.steps):
        e2e_step()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b) / args.steps

    # the e2e number is bound by the host link: report what this box's link does on a plain pinned copy of the same buffers
    def link_gbps(dst, src):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        p0, p1 = ev(), ev()
        p0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        return src.numel() * src.element_size() * 3 / (p0.elapsed_time(p1) * 1e-3) / 1e9

    link = {"h2d_GBps": round(link_gbps(x, x_host), 1), "d2h_GBps": round(link_gbps(gx_host, x), 1)}

    stats = torch.tensor([ms, e2e_ms, float(n), float(P)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, total_n, total_p = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        total_n, total_p = float(n), float(P)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, vox, pairs, threads = cpu_reference_step_time(cfg, 1, 2, 1)
            cpu = {"value": vox / sec, "unit": "voxels/s", "cores": threads, "kind": "port",
                   "sample": f"1 of {cfg['grids']} grids ({vox} voxels, {pairs} pairs) fwd+bwd fp32, median of 2 after 1 warm-up; oracle port of the GatherScatterDefault CPU path"}
        roof = dict(per_kernel[dominant])
        traffic = NCU_TRAFFIC_BYTES.get(args.config, {}).get(dominant) if cfg["grids"] == CONFIGS[args.config]["grids"] else None
        roof.update({"kernel": dominant, "traffic": traffic, "traffic_source": "ncu --set full capture, profiles/r01_ncu_full_final_summary.txt" if traffic else None,
                     "peak_source": peaks["source"]})
        line = {
            "metric": "sparse-conv voxels/sec fwd+bwd", "value": total_n / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": cfg["dtype"],
            "data": "synthetic",
            "config": {"workload": cfg["desc"], "grids_per_gpu": len(coords), "voxels_per_gpu": n, "pairs_per_gpu": P, "pairs_per_voxel": P / max(n, 1),
                       "kernel": f"{k}^3 stride 1 same-topology", "channels": f"{cin}->{cout}", "l2_policy": "inputs larger than L2 (features+grads+maps > 126 MB)",
                       "collective": "all_reduce(grad_weights), asynchronous behind wgrad, overlapping dgrad" if world > 1 else "none", "plan_build_ms": plan_ms},
            "voxel_features_per_s": total_n * (cin + cout) / 2 / (ms * 1e-3),
            "roofline": roof, "roofline_kernels": per_kernel,
            "roofline_step": {"roofline_ms": roof_time * 1e3, "measured_ms": fwd_ms + bwd_ms, "frac": roof_time * 1e3 / (fwd_ms + bwd_ms)},
            "phase_ms": {"fwd": fwd_ms, "dgrad+wgrad": bwd_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": total_n / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms, "mode": e2e_mode, "host_link_probe": link,
                    "h2d_bytes_per_step": int(x_host.numel() * s + dy_host.numel() * s), "d2h_bytes_per_step": int((y_host.numel() + gx_host.numel() + gw_host.numel()) * s)},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# C3: sparse UNet block stack training step (strided down-convs, exact-transpose up-convs, wgrad all-reduce)
# ------------------------------------------------------------------------------------------------------


def run_unet(args, cfg):
    import torch.distributed as dist

    import fvdb
    from fvdb.distributed import allreduce_gradients
    from fvdb._lib import launch_count

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = DTYPES[cfg["dtype"]]
    grids_here = cfg["grids"] // world if world <= cfg["grids"] else 1  # the batch is partitioned BY GRID (strong scaling)
    coords = make_coords({**cfg, "grids": grids_here}, rank, dev)
    g0 = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
    widths = [32, 64, 128, 256]
    Plan = fvdb.ConvolutionPlan
    grids, same, down = [g0], [], []
    for level in range(4):
        same.append(Plan.from_grid_batch(3, 1, grids[level], grids[level]))
        if level < 3:
            down.append(Plan.from_grid_batch(2, 2, grids[level]))
            grids.append(down[-1].target_grid_batch)
    up = [Plan.from_plan_transposed(p) for p in down]  # exact adjoint topology of the matching down-conv

    class Block(torch.nn.Module):
        """conv -> BatchNorm -> ReLU.  Default: fvdb.nn.BatchNorm with the ReLU fused (csrc/norm.cu); --torch-bn runs the
        reference's composition (torch.nn.BatchNorm1d over jdata, then a separate ReLU pass)."""

        def __init__(self, conv, channels):
            super().__init__()
            self.conv = conv
            if args.torch_bn:
                self.norm = torch.nn.BatchNorm1d(channels)
            else:
                self.norm = (fvdb.nn.SyncBatchNorm if args.sync_bn else fvdb.nn.BatchNorm)(channels, activation="relu")

        def forward(self, x, plan):
            y = self.conv(x, plan)
            if args.torch_bn:
                return y.jagged_like(torch.relu(self.norm(y.jdata)))
            return self.norm(y)

    class Stack(torch.nn.Module):
        def __init__(self):
            super().__init__()
            nn = fvdb.nn
            self.enc = torch.nn.ModuleList([torch.nn.ModuleList([Block(nn.SparseConv3d(c, c, 3), c) for _ in range(2)]) for c in widths])
            self.down = torch.nn.ModuleList([Block(nn.SparseConv3d(widths[i], widths[i + 1], 2, 2), widths[i + 1]) for i in range(3)])
            self.up = torch.nn.ModuleList([Block(nn.SparseConvTranspose3d(widths[i + 1], widths[i], 2, 2), widths[i]) for i in range(3)])
            self.dec = torch.nn.ModuleList([torch.nn.ModuleList([Block(nn.SparseConv3d(c, c, 3), c) for _ in range(2)]) for c in widths[:3]])

        def forward(self, x):
            skips = []
            for level in range(4):
                for blk in self.enc[level]:
                    x = blk(x, same[level])
                if level < 3:
                    skips.append(x)
                    x = self.down[level](x, down[level])
            for level in (2, 1, 0):
                x = self.up[level](x, up[level])
                x = x.jagged_like(x.jdata + skips[level].jdata)
                for blk in self.dec[level]:
                    x = blk(x, same[level])
            return x

    torch.manual_seed(1234)  # identical replicas on every rank
    model = Stack().to(dev).to(dtype)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(1 + rank)
    n = g0.total_voxels
    x_host = torch.randn((n, 32), generator=gen).to(dtype).pin_memory()
    feats = g0.jagged_like(x_host.to(dev))
    collectives = 0

    def step():
        nonlocal collectives
        opt.zero_grad(set_to_none=True)
        out = model(feats)
        loss = out.jdata.float().square().mean()
        loss.backward()
        collectives = allreduce_gradients(model.parameters()) if world > 1 else 0
        opt.step()
        return loss

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    run_step, graph_note = step, "eager launches"
    if args.graph and world == 1:  # (capturing the NCCL all-reduce with the step hung at N = 2 on this stack: eager there)
        # the whole training step (forward, backward, SGD update) as ONE CUDA graph: every kernel of the step -- ours through
        # the C ABI and torch's elementwise ones -- is captured on a side stream once and replayed per step
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l_cap = launch_count()
            with torch.cuda.graph(graph, stream=side):
                static_loss = step()
            launches_per_replay = launch_count() - l_cap

            def run_step():
                graph.replay()
                return static_loss

            graph_note = f"CUDA graph replay ({launches_per_replay} of our kernels + torch's per step in one graph launch)"
            for _ in range(2):
                run_step()
            torch.cuda.synchronize()
        except Exception as exc:  # capture is an optimisation; say so rather than hide it
            torch.cuda.synchronize()
            run_step, graph_note = step, f"eager launches (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
    l0 = launch_count()
    with ClockSampler(local_rank) as clocks:
        a, b = ev(), ev()
        a.record()
        for _ in range(args.steps):
            loss = run_step()
        b.record()
        barrier()
    launches = launch_count() - l0
    if graph_note.startswith("CUDA graph"):
        launches = launches_per_replay * args.steps
    ms = a.elapsed_time(b) / args.steps
    if args.profile and rank == 0:  # where does the step go?  (torch.profiler sees the ctypes-launched kernels through CUPTI)
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        from torch.autograd import DeviceType

        events = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA and e.device_time_total > 0]
        total = sum(e.device_time_total for e in events)
        print(f"[profile] one step: GPU kernel time {total / 1e3:.2f} ms over {sum(e.count for e in events)} launches; wall {ms:.2f} ms", file=sys.stderr)
        for e in sorted(events, key=lambda e: -e.device_time_total)[:25]:
            print(f"[profile] {e.device_time_total / 1e3:9.3f} ms {e.count:5d}  {e.key[:110]}", file=sys.stderr)
    stats = torch.tensor([ms, float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        mx, sm = stats.clone(), stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, total_n = float(mx[0]), float(sm[1])
    else:
        total_n = float(n)
    if rank == 0:
        pairs = {f"L{lv}": int(same[lv]._backend.topology.total_pairs) for lv in range(4)}
        line = {
            "metric": "sparse-conv voxels/sec fwd+bwd", "value": total_n / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["desc"], "grids_per_gpu": grids_here, "voxels_per_gpu": n, "voxels_per_level": [g.total_voxels for g in grids],
                       "pairs_3x3x3_per_level": pairs, "layers": "2x[3^3 c->c] per level, 2^3 s2 down 32-64-128-256, exact-transpose up, "
                                 + ("BN+ReLU in torch" if args.torch_bn else ("fused SyncBatchNorm+ReLU" if args.sync_bn else "fused BatchNorm+ReLU") + " (csrc/norm.cu)") + ", SGD step",
                       "collective": f"{collectives} bucketed all_reduce(grad) calls per step" if world > 1 else "none", "launch_mode": graph_note},
            "loss": float(loss), "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": None, "cpu_baseline": None, "e2e": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="row chunks of the host-buffer pipeline (upload / convolve / read back overlap)")
    ap.add_argument("--graph", action="store_true", help="c3 (1 GPU): capture the training step in one CUDA graph and replay it")
    ap.add_argument("--profile", action="store_true", help="c3: print a torch.profiler kernel-time summary of one step to stderr")
    ap.add_argument("--torch-bn", action="store_true", help="c3: torch BatchNorm1d + separate ReLU (the reference's composition) instead of the fused kernels")
    ap.add_argument("--sync-bn", action="store_true", help="c3: batch statistics over all ranks (fvdb.nn.SyncBatchNorm)")
    ap.add_argument("--grids", type=int, default=0, help="override the number of grids per GPU (experiments)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.grids:
        cfg["grids"] = args.grids
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.config == "c3":
        run_unet(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- sparse-conv voxels/sec forward+backward over a GridBatch (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c2f32|c2x128|c3|c4|c5|c5f32]

One "step" = one pass of the hot path over one batch: forward + dgrad + wgrad of a 3^3 64->64 bf16
SparseConv3d-shaped plan (BASELINE.json configs[1]: 8 indoor grids x ~200 k voxels, same-topology target),
kernel map prebuilt ("topology amortized", as the reference's own benches do).  For N > 1 (launched under
torch.distributed.run) every rank owns whole grids and the step contains the NCCL all-reduce of grad_weights, the
path's only exchange, issued asynchronously behind wgrad so that it overlaps dgrad.

The default run (no --config) prints ONE JSON line on rank 0 whose headline is C2 (weak scaling: every rank its own 8
grids, the N = 1 anchor) and which carries two sub-records measured in the same launch on the same ranks:
  "strong_c4"  BASELINE.json configs[3]: ONE batch of 32 LiDAR grids x ~1 M voxels, 3^3 128->128 bf16, partitioned BY GRID
               (LPT on voxel counts: 32 / 16 / 8 / 4 grids per GPU) -- the north_star's multi-GPU split, strong scaling;
  "train_c3"   BASELINE.json configs[2]: the sparse UNet block-stack training step on 16 indoor grids partitioned by grid,
               weight-gradient all-reduce over NCCL, replayed as one CUDA graph when capture succeeds.
`value` is timed with inputs resident in HBM; `e2e` goes through the same C-ABI-backed calls with HOST (pinned) buffers,
H2D and D2H copies inside the timed region.  `roofline` reports the dominant kernel against the gathered-bytes model of
SURVEY.md section 8(d) AND against the compulsory bytes (every row once); `gpu_baseline` is the reference's CUDA pipeline
(per-tap index_select -> mm -> index_add_, GatherScatterDefault.cu:706-721,786-808) restated in torch on the same device.
`--impl reference` times the CPU restatement of the reference's own CPU path (oracle/, torch::mm semantics, all host
threads) on the SAME config (all grids); it never loads the CUDA library.
"""

from __future__ import annotations

import argparse
import importlib.util
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
    "c2x256": dict(gen="indoor_room", grids=8, voxels=200_000, kernel=3, cin=256, cout=256, dtype="bf16", desc="C2-shaped 8 grids x ~200k voxels, 3^3 256->256 bf16 fwd+bwd"),
    "c4": dict(gen="lidar_sweep", grids=32, voxels=1_000_000, kernel=3, cin=128, cout=128, dtype="bf16", partition="by_grid",
               desc="C4 KITTI-shaped 32 grids x ~1M voxels, 3^3 128->128 bf16, GridBatch partitioned by grid"),
    "c5": dict(gen="random_occupancy", grids=8, voxels=4_979_000, kernel=5, cin=16, cout=16, dtype="bf16", desc="C5 8 grids x ~5M voxels, 5^3 16->16"),
    "c5f32": dict(gen="random_occupancy", grids=8, voxels=4_979_000, kernel=5, cin=16, cout=16, dtype="f32", desc="C5 8 grids x ~5M voxels, 5^3 16->16 fp32 (three-way bf16 split on the tensor pipe)"),
}
DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16, "f64": torch.float64}


def load_peaks() -> dict:
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(config: str) -> "dict | None":
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernels, as written by scripts/ncu_summary.py from an
    `ncu --set full` capture of this very command (profiles/r02_ncu_traffic.json: {config: {kernel: bytes}, "commit": ...})."""
    path = REPO / "profiles" / "r02_ncu_traffic.json"
    if not path.exists():
        return None
    d = json.loads(path.read_text())
    rec = d.get(config)
    return None if rec is None else {"bytes": rec, "source": f"profiles/r02_ncu_traffic.json (ncu --set full, commit {d.get('commit', '?')})"}


_synthetic = None


def synthetic():
    """The generators (pure torch) loaded by file path, so that the CPU reference arm never imports the fvdb package and
    therefore never maps libfvdbconv.so."""
    global _synthetic
    if _synthetic is None:
        spec = importlib.util.spec_from_file_location("fvdb_b200_synthetic", REPO / "fvdb-core_b200" / "fvdb" / "utils" / "synthetic.py")
        _synthetic = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_synthetic)
    return _synthetic


def partition_grids_lpt(voxel_counts, world_size):
    """Longest-processing-time bin packing by voxel count (same rule as fvdb.distributed.partition_grids_lpt, restated here so
    that the generators stay importable without the CUDA library)."""
    loads, bins = [0] * world_size, [[] for _ in range(world_size)]
    for g in sorted(range(len(voxel_counts)), key=lambda g: (-int(voxel_counts[g]), g)):
        r = min(range(world_size), key=lambda i: (loads[i], i))
        bins[r].append(g)
        loads[r] += int(voxel_counts[g])
    return [sorted(b) for b in bins]


def make_coords(cfg: dict, rank: int, device, world: int = 1) -> list[torch.Tensor]:
    """This rank's grids.  Weak scaling (default): every rank builds its own cfg["grids"] grids.  cfg["partition"] ==
    "by_grid" (C4, C3): the batch of cfg["grids"] grids is ONE job partitioned by grid -- every rank generates the whole
    batch (same seeds), bin-packs it by voxel count (LPT) and keeps its share, so no data crosses ranks."""
    gen = getattr(synthetic(), cfg["gen"])
    if cfg.get("partition") == "by_grid" and world > 1:
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
    """gathered bytes per pass (every pair charged at HBM rate)."""
    out = {
        "fwd": P * cin * s + n_out * cout * s + 4 * P + k3 * cin * cout * s,
        "dgrad": P * cout * s + n_in * cin * s + 4 * P + k3 * cin * cout * s,
        "wgrad": P * (cin + cout) * s + 8 * P + 4 * k3 * cin * cout,
    }
    # narrow layers run dgrad AND wgrad as ONE kernel off one gather of grad_output (csrc/conv_tc_bwd.cu).  The algorithmic bytes
    # are the algorithm's, not the implementation's (SURVEY.md section 8(d) charges every pass its gathered pairs): the fused
    # kernel is credited with the two passes it replaces; what it really needs is reported next to it as `fused_formulation_bytes`
    out["bwd_fused"] = out["dgrad"] + out["wgrad"]
    return out


def fused_formulation_bytes(P, n_in, n_out, cin, cout, k3, s):
    """What one gather of grad_output has to move for dgrad + wgrad together: the gathered rows and the map once, the feature
    rows read once as contiguous tiles, grad_features written once."""
    return P * cout * s + 2 * n_in * cin * s + 4 * P + k3 * cin * cout * s + 4 * k3 * cin * cout


def compulsory_bytes(P, n_in, n_out, cin, cout, k3, s):
    """SURVEY.md section 8(d) lower bound: every feature / gradient row read ONCE (P*C*s -> N*C*s), the map once."""
    return {
        "fwd": n_in * cin * s + n_out * cout * s + 4 * P + k3 * cin * cout * s,
        "dgrad": n_out * cout * s + n_in * cin * s + 4 * P + k3 * cin * cout * s,
        "wgrad": n_in * cin * s + n_out * cout * s + 4 * P + 4 * k3 * cin * cout,
        "bwd_fused": n_out * cout * s + 2 * n_in * cin * s + 4 * P + k3 * cin * cout * s + 4 * k3 * cin * cout,
    }


def kernel_rooflines(kern_ms, P, n_in, n_out, cin, cout, k3, s, peaks):
    abytes, cbytes = algorithmic_bytes(P, n_in, n_out, cin, cout, k3, s), compulsory_bytes(P, n_in, n_out, cin, cout, k3, s)
    flops = 2.0 * P * cin * cout
    out = {}
    for name, ms in kern_ms.items():
        t = ms * 1e-3
        flops = (2.0 if name != "bwd_fused" else 4.0) * P * cin * cout
        hbm_t, tensor_t = abytes[name] / (peaks["hbm_gbs"] * 1e9), flops / (peaks["tflops"] * 1e12)
        bound = "hbm" if hbm_t >= tensor_t else "tensor"
        achieved = abytes[name] / t / 1e9 if bound == "hbm" else flops / t / 1e12
        peak = peaks["hbm_gbs"] if bound == "hbm" else peaks["tflops"]
        comp_t = max(cbytes[name] / (peaks["hbm_gbs"] * 1e9), tensor_t)  # the floor no gather strategy can beat
        out[name] = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": achieved / peak, "ms": ms,
                     "algorithmic_bytes": abytes[name], "flops": flops,
                     **({"fused_formulation_bytes": fused_formulation_bytes(P, n_in, n_out, cin, cout, k3, s),
                         "note": "one kernel for dgrad + wgrad: credited with the algorithmic bytes of the two passes it replaces"} if name == "bwd_fused" else {}),
                     "compulsory": {"bytes": cbytes[name], "floor_ms": comp_t * 1e3, "frac": comp_t / t,
                                    "note": "every row once + map + weights at the HBM peak (or the FLOPs at the tensor peak, whichever is slower)"}}
    # the kernels one step launches: forward + the fused backward where it serves the layer, else forward + dgrad + wgrad
    in_step = ("fwd", "bwd_fused") if "bwd_fused" in kern_ms else ("fwd", "dgrad", "wgrad")
    roof_ms = sum(max(abytes[nm] / (peaks["hbm_gbs"] * 1e9), out[nm]["flops"] / (peaks["tflops"] * 1e12)) for nm in in_step) * 1e3
    comp_ms = sum(out[nm]["compulsory"]["floor_ms"] for nm in in_step)
    return out, roof_ms, comp_ms


# ------------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference's CPU path (never touches the CUDA library)
# ------------------------------------------------------------------------------------------------------


def oracle_problem(cfg: dict, grids: int):
    import oracle

    coords = make_coords({**cfg, "grids": grids, "partition": None}, 0, "cpu")
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
    return topo, x, w, dy


def cpu_reference_step_time(cfg: dict, sample_grids: int, repeats: int, warmup: int, budget_s: float = 150.0):
    """Seconds per fwd+bwd of the oracle (per-tap index_select -> mm -> index_add_, fp32, all host threads)
    on `sample_grids` grids of the workload.  Returns (seconds per step, voxels per step, pairs, threads)."""
    import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    topo, x, w, dy = oracle_problem(cfg, sample_grids)
    times, begin = [], time.perf_counter()
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        oracle.gs_conv(x, w, topo)
        oracle.gs_conv_backward(dy, x, w, topo)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        if times and time.perf_counter() - begin > budget_s:  # keep the whole run within a few minutes
            break
    return float(np.median(times)), int(x.shape[0]), topo.total_pairs, torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the SAME config as our arm: every grid of the batch in every step (C4 / C5 are bounded to what a CPU finishes in minutes)
    grids = cfg["grids"] if cfg["grids"] * cfg["voxels"] <= 2_000_000 else max(1, 2_000_000 // cfg["voxels"])
    sec, voxels, pairs, threads = cpu_reference_step_time(cfg, grids, max(1, args.steps), max(0, min(args.warmup, 2)))
    value = voxels / sec
    sample = f"{grids} of {cfg['grids']} grids per step ({voxels} voxels, {pairs} pairs), fp32 (the reference's CPU mm promotes half to fp32)"
    line = {
        "impl": "reference", "metric": "sparse-conv voxels/sec fwd+bwd", "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": cfg["desc"], "sample": sample, "same_config": grids == cfg["grids"]},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": sample + "; oracle port of the GatherScatterDefault CPU path (the compiled reference cannot be built offline)"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "native_library_loaded": any("libfvdbconv" in line for line in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------


class Dist:
    """One process per GPU; NCCL only where the path exchanges something (the weight-gradient all-reduce)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_host_memory_to_gpu(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.tolist()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


def bind_host_memory_to_gpu(local_rank: int) -> dict:
    """e2e at N > 1 is bound by the host copy path: keep this rank's threads and its (first-touch) pinned buffers on the NUMA
    node its GPU hangs off.  Best effort (a cpuset may forbid it); what happened is reported in the JSON line."""
    info = {"gpu_numa_node": None, "cpus_bound": None, "note": None}
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node")
        node = int(path.read_text().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            info["note"] = "no NUMA affinity reported for the device"
            return info
        cpulist = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            info["cpus_bound"] = len(target)
        else:
            info["note"] = f"cpuset {sorted(allowed)[:1]}..{sorted(allowed)[-1:]} has no CPU of node {node}: threads stay where they are"
        try:  # prefer the node for page placement even when the CPUs could not move (set_mempolicy MPOL_PREFERRED = 1)
            import ctypes

            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))  # __NR_set_mempolicy on x86_64
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
        except Exception as exc:  # noqa: BLE001
            info["mempolicy"] = f"unavailable ({type(exc).__name__})"
    except Exception as exc:  # noqa: BLE001
        info["note"] = f"NUMA binding skipped: {type(exc).__name__}: {str(exc)[:80]}"
    return info


def ev():
    return torch.cuda.Event(enable_timing=True)


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


def torch_gpu_baseline_ms(topo, x, w, dy, reps: int = 3):
    """The reference's CUDA pipeline restated in torch on this device: per tap index_select -> mm -> index_add_ forward
    (GatherScatterDefault.cu:706-721), and for the backward dgrad + wgrad per tap (:786-808).  cuBLAS + torch's own gather /
    atomic scatter kernels, same dtype as our arm."""
    k3 = topo.kernel_volume
    cout, cin = w.shape[0], w.shape[1]
    gather, scatter = topo.gather_indices.long(), topo.scatter_indices.long()
    offs = topo.offsets.tolist()
    W = w.permute(2, 3, 4, 1, 0).reshape(k3, cin, cout).contiguous()
    n_out, n_in = topo.output_total_voxels, topo.feature_total_voxels

    def step():
        y = torch.zeros((n_out, cout), dtype=x.dtype, device=x.device)
        for k in range(k3):
            a, b = offs[k], offs[k + 1]
            if b > a:
                y.index_add_(0, scatter[a:b], x.index_select(0, gather[a:b]) @ W[k])
        gx = torch.zeros((n_in, cin), dtype=x.dtype, device=x.device)
        gw = torch.zeros_like(W)
        for k in range(k3):
            a, b = offs[k], offs[k + 1]
            if b > a:
                g = dy.index_select(0, scatter[a:b])
                gx.index_add_(0, gather[a:b], g @ W[k].T)
                gw[k] = x.index_select(0, gather[a:b]).T @ g
        return y, gx, gw

    return timed(step, reps)


def conv_workload(D: Dist, cfg: dict, args, *, want_e2e: bool, want_gpu_baseline: bool, steps: int, warmup: int):
    """Build this rank's share of `cfg`, time fwd + bwd (+ all-reduce) and the three hot kernels; returns a result dict."""
    import torch.distributed as dist

    import fvdb
    from fvdb import _fvdb_cpp as cpp
    from fvdb._lib import launch_count

    world, dev = D.world, D.dev
    dtype = DTYPES[cfg["dtype"]]
    k, cin, cout = cfg["kernel"], cfg["cin"], cfg["cout"]
    coords = make_coords(cfg, D.rank, dev, world)
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
    del coords
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid)
    topo = plan._backend.topology
    topo._dgrad_plan()  # the input-stationary map for dgrad is part of the (amortised) plan (a symmetric map serves as it is)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    n, P, k3 = grid.total_voxels, topo.total_pairs, topo.kernel_volume

    gen = torch.Generator().manual_seed(1 + D.rank)
    if want_e2e:
        x_host = torch.randn((n, cin), generator=gen).to(dtype).pin_memory()
        dy_host = torch.randn((n, cout), generator=gen).to(dtype).pin_memory()
        x, dy = x_host.to(dev), dy_host.to(dev)
    else:  # big strong-scaling batches: generate on the device
        dgen = torch.Generator(device=dev).manual_seed(1 + D.rank)
        x = torch.randn((n, cin), generator=dgen, device=dev).to(dtype)
        dy = torch.randn((n, cout), generator=dgen, device=dev).to(dtype)
    w = ((torch.rand((cout, cin, k, k, k), generator=gen) * 2 - 1) / (cin * k**3) ** 0.5).to(dtype).to(dev)

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

    if world > 1:  # untimed set-up: NCCL builds its channels lazily on the first collectives of a given size
        probe = torch.zeros_like(w)
        for _ in range(8):
            dist.all_reduce(probe)
        torch.cuda.synchronize()
    for _ in range(warmup):
        step(False)
    D.barrier()
    launches0 = launch_count()
    with ClockSampler(D.local_rank) as clocks:
        start, stop = ev(), ev()
        start.record()
        for _ in range(steps):
            step(True)
        stop.record()
        D.barrier()
    launches = launch_count() - launches0
    ms = start.elapsed_time(stop) / steps
    fwd_ms = float(np.mean([a.elapsed_time(b) for a, b in phase_ms["fwd"]]))
    bwd_ms = float(np.mean([a.elapsed_time(b) for a, b in phase_ms["dgrad+wgrad"]]))
    # per-step device times (median and min next to the mean the contract's ms_per_step is)
    per_step = [f[0].elapsed_time(b[1]) for f, b in zip(phase_ms["fwd"], phase_ms["dgrad+wgrad"])]

    # the three hot kernels, each timed alone on the launching stream (weights prepared outside the loop, as a layer would cache them)
    w_fwd = cpp._prepare_weights(w, dtype, False)
    in_map, in_mask, mirror = topo._dgrad_plan()
    w_bwd = cpp._prepare_weights(w, dtype, True, flip_taps=mirror)
    out_map = topo._out_map()
    reps = max(3, steps)
    kern_ms = {
        "fwd": timed(lambda: cpp._run_conv(x, w_fwd, out_map, n, n, cin, cout, k3, None, topo._out_mask()), reps),
        "dgrad": timed(lambda: cpp._run_conv(dy, w_bwd, in_map, n, n, cout, cin, k3, None, in_mask), reps),
        "wgrad": timed(lambda: cpp.gs_conv_backward(dy, x, w, topo, need_grad_features=False), reps),
    }
    if cpp._fused_backward and dtype in (torch.float16, torch.bfloat16) and int(cpp.lib.fvc_conv_kernel_family(cin, cout, k3, cpp._DTYPE_CODE[dtype], 0, 2)) == 2:
        kern_ms["bwd_fused"] = timed(lambda: cpp.gs_conv_backward(dy, x, w, topo), reps)  # (its weight image and partial reduction included)
    peaks = load_peaks()
    s = x.element_size()
    per_kernel, roof_ms, comp_ms = kernel_rooflines(kern_ms, P, n, n, cin, cout, k3, s, peaks)

    res = {"ms": ms, "n": n, "P": P, "k3": k3, "grids": grid.grid_count, "plan_ms": plan_ms, "launches": launches, "clocks": clocks.summary(),
           "per_kernel": per_kernel, "roof_ms": roof_ms, "comp_ms": comp_ms, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "peaks": peaks, "elem": s,
           "step_median_ms": float(np.median(per_step)), "step_min_ms": float(np.min(per_step)),
           "gpu_baseline_ms": None, "e2e": None}
    if want_gpu_baseline:
        try:
            res["gpu_baseline_ms"] = torch_gpu_baseline_ms(topo, x, w, dy)
        except RuntimeError as exc:  # out of memory on a very large config: say so rather than fail the bench
            res["gpu_baseline_error"] = str(exc)[:120]
            torch.cuda.empty_cache()
    if want_e2e:
        res["e2e"] = e2e_through_host_buffers(D, args, plan, topo, x, x_host, dy_host, w, n, cin, cout, k3, dtype, steps, warmup)
    return res


def e2e_through_host_buffers(D, args, plan, topo, x, x_host, dy_host, w, n, cin, cout, k3, dtype, steps, warmup):
    """End to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region."""
    import torch.distributed as dist

    from fvdb import _fvdb_cpp as cpp

    world, dev = D.world, D.dev
    s = x.element_size()
    code = cpp._DTYPE_CODE[dtype]
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
    if cpp.lib.fvc_conv_scratch_bytes(n, n, cin, cout, k3, code) > 0 and dtype != torch.float32 and bool(plan.has_fixed_topology):  # tensor-core path available
        from fvdb.streaming import HostPipelinedConv

        pipe = HostPipelinedConv(plan, num_chunks=args.e2e_chunks)
        e2e_mode = f"3-stream overlap, {len(pipe.bounds)} row chunks (fvdb.streaming.HostPipelinedConv)"

        def e2e_step():  # noqa: F811
            pipe.forward_backward(x_host, dy_host, w, y_host, gx_host, gw_host, reduce_fn=(lambda g: dist.all_reduce(g)) if world > 1 else None)

    for _ in range(max(1, min(warmup, 3))):
        e2e_step()
    D.barrier()
    a, b = ev(), ev()
    a.record()
    for _ in range(steps):
        e2e_step()
    b.record()
    D.barrier()
    e2e_ms = a.elapsed_time(b) / steps

    # the e2e number is bound by the host link: report what this rank's link does on a plain pinned copy of the same buffers,
    # with every rank copying at the same time (as in the step)
    def link_gbps(dst, src):
        dst.copy_(src, non_blocking=True)
        D.barrier()
        p0, p1 = ev(), ev()
        p0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        return src.numel() * src.element_size() * 3 / (p0.elapsed_time(p1) * 1e-3) / 1e9

    h2d, d2h = link_gbps(x, x_host), link_gbps(gx_host, x)
    return {"ms": e2e_ms, "mode": e2e_mode, "h2d_GBps": h2d, "d2h_GBps": d2h,
            "h2d_bytes": int(x_host.numel() * s + dy_host.numel() * s), "d2h_bytes": int((y_host.numel() + gx_host.numel() + gw_host.numel()) * s)}


def conv_record(D: Dist, cfg: dict, res: dict, *, steps: int, warmup: int, strong: bool, config_name: str) -> dict:
    """Reduce one workload's per-rank numbers over the ranks (MAX of times, SUM of units) -> the record rank 0 prints."""
    k, cin, cout = cfg["kernel"], cfg["cin"], cfg["cout"]
    e2e = res["e2e"]
    mx = D.reduce([res["ms"], e2e["ms"] if e2e else 0.0, res["fwd_ms"], res["bwd_ms"]] + [res["per_kernel"][nm]["ms"] for nm in ("fwd", "dgrad", "wgrad")], "max")
    sm = D.reduce([float(res["n"]), float(res["P"]), e2e["h2d_GBps"] if e2e else 0.0, e2e["d2h_GBps"] if e2e else 0.0], "sum")
    mn = D.reduce([-(e2e["h2d_GBps"] if e2e else 0.0), -(e2e["d2h_GBps"] if e2e else 0.0)], "max")
    ms, e2e_ms = mx[0], mx[1]
    total_n, total_p = sm[0], sm[1]
    per_kernel = res["per_kernel"]
    in_step = ("fwd", "bwd_fused") if "bwd_fused" in per_kernel else ("fwd", "dgrad", "wgrad")  # the kernels a step launches
    dominant = max(in_step, key=lambda nm: per_kernel[nm]["ms"])
    roof = dict(per_kernel[dominant])
    traffic = ncu_traffic(config_name) if D.world == 1 else None
    roof.update({"kernel": dominant, "traffic": (traffic["bytes"].get(dominant) if traffic else None), "traffic_source": traffic["source"] if traffic else None,
                 "peak_source": res["peaks"]["source"], "scope": "rank 0's share of the batch"})
    measured = res["fwd_ms"] + res["bwd_ms"]
    rec = {
        "metric": "sparse-conv voxels/sec fwd+bwd", "value": total_n / (ms * 1e-3), "unit": "voxels/s", "n_gpus": D.world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": cfg["desc"], "grids_per_gpu": res["grids"], "voxels_per_gpu": res["n"], "pairs_per_gpu": res["P"], "pairs_per_voxel": res["P"] / max(res["n"], 1),
                   "total_voxels": int(total_n), "total_pairs": int(total_p), "kernel": f"{k}^3 stride 1 same-topology", "channels": f"{cin}->{cout}",
                   "l2_policy": "inputs larger than L2 (features+grads+maps > 126 MB)",
                   "partition": ("ONE batch partitioned by grid (LPT on voxel counts), no data-path collective" if strong else "every rank owns its own batch of whole grids"),
                   "collective": "all_reduce(grad_weights), asynchronous behind wgrad, overlapping dgrad" if D.world > 1 else "none", "plan_build_ms": res["plan_ms"]},
        "voxel_features_per_s": total_n * (cin + cout) / 2 / (ms * 1e-3),
        "roofline": roof, "roofline_kernels": per_kernel,
        "roofline_step": {"roofline_ms": res["roof_ms"], "compulsory_ms": res["comp_ms"], "measured_ms": measured, "frac": res["roof_ms"] / measured,
                          "compulsory_frac": res["comp_ms"] / measured},
        "phase_ms": {"fwd": mx[2], "dgrad+wgrad": mx[3]},
        "step_ms": {"mean": ms, "median": res["step_median_ms"], "min": res["step_min_ms"], "note": "this rank's per-step CUDA-event times; ms_per_step is the mean over the timed region, max over ranks"},
        "gpu_launches": int(res["launches"]), "clocks": res["clocks"],
    }
    if res.get("gpu_baseline_ms"):
        rec["gpu_baseline"] = {"value": res["n"] / (res["gpu_baseline_ms"] * 1e-3), "unit": "voxels/s", "ms_per_step": res["gpu_baseline_ms"], "kind": "port",
                               "what": "the reference's CUDA pipeline (per tap index_select -> mm -> index_add_, GatherScatterDefault.cu:706-721,786-808) restated in torch "
                                       "(cuBLAS + torch gather / atomic scatter kernels) on the same device, same batch and dtype; rank 0's share"}
    if e2e:
        rec["e2e"] = {"value": total_n / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms, "mode": e2e["mode"],
                      "host_link_probe": {"h2d_GBps_sum_over_ranks": round(sm[2], 1), "d2h_GBps_sum_over_ranks": round(sm[3], 1),
                                          "h2d_GBps_slowest_rank": round(-mn[0], 1), "d2h_GBps_slowest_rank": round(-mn[1], 1),
                                          "note": "plain pinned copies of the step's buffers, all ranks at once: the aggregate host bandwidth that bounds e2e"},
                      "host_numa": D.numa, "h2d_bytes_per_step": e2e["h2d_bytes"], "d2h_bytes_per_step": e2e["d2h_bytes"]}
    return rec


def run_ours(args, cfg, config_name: str):
    D = Dist()
    strong = cfg.get("partition") == "by_grid"
    big = cfg["grids"] * cfg["voxels"] > 12_000_000  # C4 / C5: no host-buffer arm, no torch baseline (memory)
    res = conv_workload(D, cfg, args, want_e2e=not big, want_gpu_baseline=(D.world == 1 and not big) or args.gpu_baseline, steps=args.steps, warmup=args.warmup)
    line = conv_record(D, cfg, res, steps=args.steps, warmup=args.warmup, strong=strong, config_name=config_name)
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        sample = cfg["grids"] if cfg["grids"] * cfg["voxels"] <= 2_000_000 else 1
        sec, vox, pairs, threads = cpu_reference_step_time(cfg, sample, 2, 1, budget_s=60.0)
        line["cpu_baseline"] = {"value": vox / sec, "unit": "voxels/s", "cores": threads, "kind": "port",
                                "sample": f"{sample} of {cfg['grids']} grids ({vox} voxels, {pairs} pairs) fwd+bwd fp32, median of 2 after 1 warm-up; oracle port of the GatherScatterDefault CPU path"}
    else:
        line["cpu_baseline"] = None
    if args.sub_records:
        del res
        torch.cuda.empty_cache()
        sub_steps, sub_warm = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
        c4 = dict(CONFIGS["c4"])
        if args.c4_grids:
            c4["grids"] = args.c4_grids
        try:
            r4 = conv_workload(D, c4, args, want_e2e=False, want_gpu_baseline=False, steps=sub_steps, warmup=sub_warm)
            line["strong_c4"] = conv_record(D, c4, r4, steps=sub_steps, warmup=sub_warm, strong=True, config_name="c4")
            del r4
        except RuntimeError as exc:
            line["strong_c4"] = {"error": str(exc)[:200]}
        torch.cuda.empty_cache()
        try:
            line["train_c3"] = unet_record(D, args, dict(CONFIGS["c3"]), steps=sub_steps, warmup=sub_warm, graph=not args.no_graph)
        except RuntimeError as exc:
            line["train_c3"] = {"error": str(exc)[:200]}
    if D.rank == 0:
        print(json.dumps(line), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------------------
# C3: sparse UNet block stack training step (strided down-convs, exact-transpose up-convs, wgrad all-reduce)
# ------------------------------------------------------------------------------------------------------


def unet_record(D: Dist, args, cfg, *, steps: int, warmup: int, graph: bool) -> dict:
    import torch.distributed as dist

    import fvdb
    from fvdb._lib import launch_count
    from fvdb.distributed import allreduce_gradients

    world, rank, dev = D.world, D.rank, D.dev
    dtype = DTYPES[cfg["dtype"]]
    # ONE batch of cfg["grids"] indoor grids partitioned BY GRID (strong scaling); same seeds on every rank
    coords = make_coords({**cfg, "partition": "by_grid"}, rank, dev, world)
    g0 = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
    grids_here = g0.grid_count
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
        """conv -> BatchNorm -> ReLU.  Default: fvdb.nn.ConvBNReLU-style fusion (statistics from the conv epilogue, one
        apply pass); --torch-bn runs the reference's composition (torch.nn.BatchNorm1d over jdata, then a separate ReLU)."""

        def __init__(self, conv, channels):
            super().__init__()
            self.conv = conv
            if args.torch_bn:
                self.norm = torch.nn.BatchNorm1d(channels)
            else:
                self.norm = (fvdb.nn.SyncBatchNorm if args.sync_bn else fvdb.nn.BatchNorm)(channels, activation="relu")

        def forward(self, x, plan):
            if args.torch_bn:
                y = self.conv(x, plan)
                return y.jagged_like(torch.relu(self.norm(y.jdata)))
            if hasattr(fvdb.nn, "conv_bn_act") and not args.unfused_bn:
                return fvdb.nn.conv_bn_act(self.conv, self.norm, x, plan)
            return self.norm(self.conv(x, plan))

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
    params = list(model.parameters())

    def step():
        nonlocal collectives
        opt.zero_grad(set_to_none=True)
        out = model(feats)
        loss = out.jdata.float().square().mean()
        loss.backward()
        collectives = allreduce_gradients(params) if world > 1 else 0
        opt.step()
        return loss.detach()

    for _ in range(warmup):
        step()
    D.barrier()
    run_step, graph_note, launches_per_replay = step, "eager launches", None
    if graph:
        # the whole training step (forward, backward, all-reduce, SGD update) as ONE CUDA graph: every kernel of the step --
        # ours through the C ABI, torch's elementwise ones and NCCL's -- is captured on a side stream once and replayed per
        # step, so the ~600 launches of a step stop bounding it (at N > 1 each rank has 1/N of the voxels but the same launches)
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            D.barrier()
            g = torch.cuda.CUDAGraph()
            l_cap = launch_count()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                static_loss = step()
            launches_per_replay = launch_count() - l_cap

            def run_step():
                g.replay()
                return static_loss

            graph_note = f"CUDA graph replay ({launches_per_replay} of our kernels + torch's" + (" + the NCCL all-reduces" if world > 1 else "") + " per step in one graph launch)"
            for _ in range(2):
                run_step()
            D.barrier()
        except Exception as exc:  # capture is an optimisation; say so rather than hide it  # noqa: BLE001
            torch.cuda.synchronize()
            run_step, graph_note = step, f"eager launches (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
    # every rank must take the same path through the collectives: agree on graph vs eager
    if world > 1:
        flag = D.reduce([0.0 if graph_note.startswith("CUDA graph") else 1.0], "max")[0]
        if flag > 0 and graph_note.startswith("CUDA graph"):
            run_step, graph_note = step, "eager launches (another rank failed to capture)"
    l0 = launch_count()
    with ClockSampler(D.local_rank) as clocks:
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            loss = run_step()
        b.record()
        D.barrier()
    launches = launch_count() - l0
    if graph_note.startswith("CUDA graph"):
        launches = launches_per_replay * steps
    ms = a.elapsed_time(b) / steps
    # end to end: the step's input features come from pinned host memory every step and the loss is read back to the host, both
    # inside the timed region (same stream as the step: upload, step, read-back in series)
    x_dev, loss_host = feats.jdata, torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        x_dev.copy_(x_host, non_blocking=True)
        loss_host.copy_(run_step().float(), non_blocking=True)

    for _ in range(2):
        e2e_step()
    D.barrier()
    a2, b2 = ev(), ev()
    a2.record()
    for _ in range(steps):
        e2e_step()
    b2.record()
    D.barrier()
    e2e_ms = a2.elapsed_time(b2) / steps
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
    # roofline of the step: the convolution layers' gathered / compulsory bytes (fwd + dgrad + wgrad each); BN, ReLU, skips excluded
    peaks = load_peaks()
    layers = []
    for level in range(4):
        c = widths[level]
        t = same[level]._backend.topology
        layers += [(t.total_pairs, t.feature_total_voxels, t.output_total_voxels, c, c, 27)] * (2 if level == 3 else 4)
        if level < 3:
            t = down[level]._backend.topology
            layers.append((t.total_pairs, t.feature_total_voxels, t.output_total_voxels, c, widths[level + 1], 8))
            layers.append((t.total_pairs, t.output_total_voxels, t.feature_total_voxels, widths[level + 1], c, 8))  # the transposed up-conv
    roof_s = comp_s = 0.0
    for P, n_in, n_out, ci, co, k3 in layers:
        ab, cb = algorithmic_bytes(P, n_in, n_out, ci, co, k3, 2), compulsory_bytes(P, n_in, n_out, ci, co, k3, 2)
        fl = 2.0 * P * ci * co / (peaks["tflops"] * 1e12)
        roof_s += sum(max(ab[nm] / (peaks["hbm_gbs"] * 1e9), fl) for nm in ("fwd", "dgrad", "wgrad"))
        comp_s += sum(max(cb[nm] / (peaks["hbm_gbs"] * 1e9), fl) for nm in ("fwd", "dgrad", "wgrad"))
    mx = D.reduce([ms, roof_s, comp_s, e2e_ms], "max")
    total_n = D.reduce([float(n)], "sum")[0]
    ms, e2e_ms = mx[0], mx[3]
    rec = {
        "metric": "sparse-conv voxels/sec fwd+bwd", "value": total_n / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": cfg["desc"], "grids_per_gpu": grids_here, "voxels_per_gpu": n, "total_voxels": int(total_n), "voxels_per_level": [g.total_voxels for g in grids],
                   "pairs_3x3x3_per_level": {f"L{lv}": int(same[lv]._backend.topology.total_pairs) for lv in range(4)},
                   "layers": "2x[3^3 c->c] per level, 2^3 s2 down 32-64-128-256, exact-transpose up, "
                             + ("BN+ReLU in torch" if args.torch_bn else ("SyncBatchNorm" if args.sync_bn else "BatchNorm") + "+ReLU: statistics from the conv epilogue, one fused apply pass (csrc/norm.cu)") + ", SGD step",
                   "collective": f"{collectives} bucketed all_reduce(grad) calls per step (NCCL)" if world > 1 else "none", "launch_mode": graph_note},
        "loss": float(loss), "gpu_launches": int(launches), "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "unit": "ms", "scope": "the step's convolution layers (fwd + dgrad + wgrad each), slowest rank; BN / ReLU / skip adds / optimizer excluded",
                     "roofline_ms": mx[1] * 1e3, "compulsory_ms": mx[2] * 1e3, "measured_ms": ms, "frac": mx[1] * 1e3 / ms, "compulsory_frac": mx[2] * 1e3 / ms,
                     "peak": peaks["hbm_gbs"], "peak_source": peaks["source"]},
        "e2e": {"value": total_n / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size()),
                "d2h_bytes_per_step": 4, "mode": "per step: features uploaded from pinned host memory, the training step, the loss read back (in series on one stream)"},
    }
    return rec


def run_unet(args, cfg):
    D = Dist()
    rec = unet_record(D, args, cfg, steps=args.steps, warmup=args.warmup, graph=args.graph)
    if D.rank == 0:
        if D.world == 1 and not args.no_cpu_baseline:  # the stack's convolution layers through the oracle port on one of the grids
            rec["cpu_baseline"] = unet_cpu_baseline(cfg)
        print(json.dumps(rec), flush=True)
    D.close()


def unet_cpu_baseline(cfg) -> dict:
    import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    coords = make_coords({**cfg, "grids": 1, "partition": None}, 0, "cpu")
    ijk = coords[0].numpy().astype(np.int64)
    b = np.zeros(len(ijk), dtype=np.int64)
    order = oracle.index_grid_row_order(b, ijk)
    level_ijk = [ijk[order]]
    widths, total, gen = [32, 64, 128, 256], 0.0, torch.Generator().manual_seed(1)

    def conv_time(topo, ci, co, k):
        x = torch.randn((topo.feature_total_voxels, ci), generator=gen)
        w = torch.randn((co, ci, k, k, k), generator=gen) * 0.05
        dy = torch.randn((topo.output_total_voxels, co), generator=gen)
        t0 = time.perf_counter()
        oracle.gs_conv(x, w, topo)
        oracle.gs_conv_backward(dy, x, w, topo)
        return time.perf_counter() - t0

    for level in range(4):
        cur = level_ijk[level]
        zb = np.zeros(len(cur), dtype=np.int64)
        topo = oracle.build_topology(cur, zb, cur, zb, 3, 1)
        total += conv_time(topo, widths[level], widths[level], 3) * (2 if level == 3 else 4)
        if level < 3:
            c_ijk, c_b = oracle.conv_grid(cur, zb, 2, 2)
            o = oracle.index_grid_row_order(c_b, c_ijk)
            c_ijk = c_ijk[o]
            down = oracle.build_topology(cur, zb, c_ijk, np.zeros(len(c_ijk), dtype=np.int64), 2, 2)
            total += conv_time(down, widths[level], widths[level + 1], 2)
            total += conv_time(oracle.reverse_topology(down), widths[level + 1], widths[level], 2)
            level_ijk.append(c_ijk)
    return {"value": len(ijk) / total, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"the stack's 20 convolution layers (fwd + dgrad + wgrad, fp32) on 1 of {cfg['grids']} grids ({len(ijk)} voxels), one pass; oracle port of the GatherScatterDefault CPU path; BN / ReLU excluded"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="default run: skip the strong_c4 / train_c3 sub-records")
    ap.add_argument("--gpu-baseline", action="store_true", help="also time the torch restatement of the reference's CUDA pipeline where it is skipped by default")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="row chunks of the host-buffer pipeline (upload / convolve / read back overlap)")
    ap.add_argument("--graph", action="store_true", help="--config c3: capture the training step in one CUDA graph and replay it")
    ap.add_argument("--no-graph", action="store_true", help="default run: keep the train_c3 sub-record eager")
    ap.add_argument("--profile", action="store_true", help="c3: print a torch.profiler kernel-time summary of one step to stderr")
    ap.add_argument("--torch-bn", action="store_true", help="c3: torch BatchNorm1d + separate ReLU (the reference's composition) instead of the fused kernels")
    ap.add_argument("--unfused-bn", action="store_true", help="c3: fvdb.nn.BatchNorm as a separate module call (statistics pass + apply pass) instead of the conv-epilogue statistics")
    ap.add_argument("--sync-bn", action="store_true", help="c3: batch statistics over all ranks (fvdb.nn.SyncBatchNorm)")
    ap.add_argument("--grids", type=int, default=0, help="override the number of grids per GPU (experiments)")
    ap.add_argument("--variant", type=int, default=0, help="experiments: pipeline-shape variant of the tensor-core forward kernel (fvc_set_tuning)")
    ap.add_argument("--wgrad-variant", type=int, default=0, help="experiments: variant of the weight-gradient kernel (fvc_set_tuning key 1)")
    ap.add_argument("--c4-grids", type=int, default=0, help="override the number of grids of the strong_c4 sub-record (experiments)")
    args = ap.parse_args()
    config_name = args.config or "c2"
    args.sub_records = args.config is None and not args.no_sub_records and args.impl == "ours"
    cfg = dict(CONFIGS[config_name])
    if args.grids:
        cfg["grids"] = args.grids
    if args.variant and args.impl == "ours":
        from fvdb import _fvdb_cpp

        _fvdb_cpp.set_kernel_variant(args.variant)
    if args.wgrad_variant and args.impl == "ours":
        from fvdb import _fvdb_cpp

        _fvdb_cpp.set_kernel_variant(args.wgrad_variant, wgrad=True)
    if args.impl == "reference":
        run_reference(args, cfg)
    elif config_name == "c3":
        run_unet(args, cfg)
    else:
        run_ours(args, cfg, config_name)


if __name__ == "__main__":
    main()

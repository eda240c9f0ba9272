"""Benchmark of the COTR correspondence hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], the one `metric` is quoted on): per GPU one synthetic 256x256 image pair laid side
by side (1,3,256,512) and 1024 random queries, no zoom; a "step" is one pass of the hot path over that batch
(COTR.forward: backbone -> input_proj -> encoder -> decoder -> head).  Weak scaling: every rank processes its own pair
(independent pairs shard with no data-path collective); for N > 1 the (N,1024,2) results are all-gathered over NCCL
inside the timed step (the path's only exchange).

One JSON line on rank 0:
  value      query-points/s, whole job, inputs resident in HBM, CUDA-event timed per step, L2 flushed between steps
  e2e        same metric through the C-ABI host-buffer call (cotr_forward_host): pinned host buffers, H2D + D2H inside
  roofline   dominant kernel (by time) from a separate pass with per-launch CUDA events on the launching stream
  cpu_baseline  the oracle port (CPU restatement of the reference, oracle/cotr_oracle.py) on the host cores, rank 0, N=1
`--impl reference` times that CPU path alone (all host threads) and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_QUERIES = 1024
METRIC = "query-points/sec (256x256 pair, 1024 queries)"
UNIT = "query-points/s"
WORKLOAD = "configs[1]: single 256x256 synthetic pair, 1024 random queries, no zoom (per GPU)"
CTX_FLOP = 24.641536e9          # per pair
QUERY_FLOP = 11273216.0         # per query


def algorithmic_flop(pairs, queries):
    return pairs * (CTX_FLOP + queries * QUERY_FLOP)


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p["bf16_tflops"]), "hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, burst)"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """CPU threads this process may really use: min(os.cpu_count, sched affinity, cgroup cpu quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                quota = int(txt[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0:
                    n = min(n, max(1, quota // period))
            break
        except Exception:
            continue
    return max(1, n)


def cpu_reference_rate(budget_s, warmup=1, max_iters=30):
    """The oracle port (validated bit-exact against the reference on CPU) on all host threads, same workload."""
    import torch
    from oracle import cotr_oracle, fixtures
    torch.set_num_threads(host_threads())
    sd = cotr_oracle.cast_state_dict(fixtures.make_state_dict(0), torch.float32)
    img, queries = fixtures.make_inputs(1, 1, N_QUERIES)
    for _ in range(warmup):
        cotr_oracle.forward(sd, img, queries, torch.float32)
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_start < budget_s or len(times) < 3):
        t0 = time.perf_counter()
        cotr_oracle.forward(sd, img, queries, torch.float32)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": N_QUERIES / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} forwards of the same workload (1 pair, 1024 queries), median {med * 1e3:.1f} ms, "
                      f"fp32 eager torch {torch.__version__} on {torch.get_num_threads()} threads (os.cpu_count {os.cpu_count()})"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port), rank 0 only."""
    if rank != 0:
        return
    steps = []
    base = None
    import torch
    from oracle import cotr_oracle, fixtures
    torch.set_num_threads(host_threads())
    sd = cotr_oracle.cast_state_dict(fixtures.make_state_dict(0), torch.float32)
    img, queries = fixtures.make_inputs(1, 1, N_QUERIES)
    for _ in range(max(args.warmup, 1)):
        cotr_oracle.forward(sd, img, queries, torch.float32)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cotr_oracle.forward(sd, img, queries, torch.float32)
        steps.append(time.perf_counter() - t0)
    ms = float(np.mean(steps)) * 1e3
    value = N_QUERIES / (ms * 1e-3)
    base = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{args.steps} forwards of 1 pair x 1024 queries per step (the whole workload of one GPU), fp32 eager torch on {torch.get_num_threads()} threads (os.cpu_count {os.cpu_count()})"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU path does not use the GPUs; one pair per step"},
        "cpu_baseline": base, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from cotr_b200.models import build_model
    from cotr_b200.utils import synthetic as fixtures   # seeded synthetic weights / inputs (nothing under oracle/ on this path)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sd = fixtures.make_state_dict(0)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(dev).eval()
    img_np, q_np = fixtures.make_inputs(100 + rank, 1, N_QUERIES)
    img = torch.from_numpy(img_np).to(dev)
    queries = torch.from_numpy(q_np).to(dev)
    img_pin = torch.from_numpy(img_np).pin_memory()
    q_pin = torch.from_numpy(q_np).pin_memory()
    out_pin = torch.empty((1, N_QUERIES, 2), dtype=torch.float32).pin_memory()
    gathered = torch.empty((world, N_QUERIES, 2), dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step():
        pred = model(img, queries)["pred_corrs"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, pred)
        return pred

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nat = model.native()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps ----------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        step()
        stops[i].record()
    barrier()
    launches_per_step = nat.last_launch_count()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / args.steps

    # ---- e2e: the C-ABI host-buffer call, pinned host memory, H2D + forward + D2H + sync inside ---------------------
    img_h, q_h, out_h = img_pin.numpy(), q_pin.numpy(), out_pin.numpy()
    for _ in range(3):
        nat.forward_host(img_h, q_h, out_h)
    barrier()
    e2e_times = []
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nat.forward_host(img_h, q_h, out_h)
        e2e_times.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline pass: per-launch events (library profiler), same steps ------------------------------------------
    per_kernel = {}
    if rank == 0:
        for _ in range(args.steps):
            flush.zero_()
            nat.profile_begin(1024)
            model(img, queries)
            for name, M, N, K, ms in nat.profile_end():
                d = per_kernel.setdefault(name, {"ms": 0.0, "launches": 0, "flop": 0.0})
                d["ms"] += ms; d["launches"] += 1
                if name.startswith("gemm"):
                    d["flop"] += 2.0 * M * N * K
                elif name.startswith("attention"):
                    d["flop"] += 2.0 * 2.0 * M * N * K        # QK^T + PV over 8 heads x 32 dims
    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    if rank != 0:
        return

    peaks = measured_peaks()
    total_q = world * N_QUERIES
    value = total_q / (dev_ms * 1e-3)
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
    d = per_kernel[dom]
    achieved_tf = d["flop"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    step_tf = algorithmic_flop(1, N_QUERIES) / (dev_ms * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved_tf / peaks["bf16_tflops"],
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the four gemm_tc launches of the committed
        # `ncu --set full` capture (profiles/r01_final_ncu_summary.md: 2.93 / 1.36 / 1.63 / 3.72 MB read, 0 written;
        # algorithmic bytes of the same launches 2.8 / 1.3 / 1.6 / 3.4 MB)
        "traffic": 2.41e6 if dom == "gemm_tc" else None, "traffic_unit": "bytes per launch (ncu capture, not measured live)",
        "peak_source": peaks["source"],
        "launches_per_step": d["launches"] // args.steps, "kernel_ms_per_step": d["ms"] / args.steps,
        "kernel_share_of_step": d["ms"] / sum(v["ms"] for v in per_kernel.values()),
        "whole_step": {"algorithmic_gflop": algorithmic_flop(1, N_QUERIES) / 1e9, "achieved": step_tf, "frac": step_tf / peaks["bf16_tflops"]},
        "per_kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(per_kernel.items())},
        "note": "algorithmic FLOPs (2*M*N*K per GEMM launch; the 3 split-precision MMAs per product are NOT counted) / summed launch durations; "
                "measured in a separate pass with per-launch CUDA events on the launching stream",
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 hi/lo split operands, fp32 accumulate on tcgen05)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": 1, "queries_per_pair": N_QUERIES, "parallelism": f"dp{world} (independent pairs)",
                   "l2": "flushed between timed steps by writing a 256 MiB buffer", "weights": "seeded synthetic (cotr_b200/utils/synthetic.py seed 0)",
                   "result_gather": "nccl all_gather inside the step" if world > 1 else "none (single GPU)"},
        "e2e": {"value": total_q / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(img_np.nbytes + q_np.nbytes), "d2h_bytes_per_step": int(out_pin.numel() * 4),
                "api": "cotr_forward_host (C ABI, pinned host buffers)"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "clocks": clocks,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_reference_rate(budget_s=10.0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_native(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

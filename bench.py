#!/usr/bin/env python
"""bench.py -- TGN train events/sec (+ eval users/sec) of the PfoTGNRec hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores (oracle port)

A step = one optimiser step over one batch of the synthetic stream: candidate sampling + MV
selection, neighbour sampling, lazy memory update, temporal attention, BPR, backward, Adam,
memory persist + message store (reference main.py:179-394).  One JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ours", choices=["ours", "tgn", "jodie", "dyrep", "tgat"])
    ap.add_argument("--bs", type=int, default=8192)
    ap.add_argument("--users", type=int, default=100000)
    ap.add_argument("--items", type=int, default=1000)
    ap.add_argument("--events", type=int, default=5000000)
    ap.add_argument("--days", type=int, default=200)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--eval-steps", type=int, default=4)
    ap.add_argument("--eval-bs", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"{'PfoTGNRec' if a.workload == 'ours' else a.workload} train step "
            f"(d=64, 1 layer, 10 neighbours, 2 heads{', MV sampling K=20' if a.workload == 'ours' else ''}), "
            f"synthetic {a.users}-user x {a.items}-stock x {a.events}-event stream, bs={a.bs}")


def make_data(a):
    from pfotgnrec_b200.synth import make_stream
    return make_stream(n_users=a.users, n_items=a.items, n_events=a.events, n_days=a.days, seed=0, ts_mode="nbg")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(a):
    """The reference algorithm on the host cores: the oracle port (oracle/train_loop.py).  The reference
    itself is pure Python and cannot travel to the GPU box (no /root/reference there); see DESIGN.md."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.train_loop import OracleTrainer
    torch.set_num_threads(os.cpu_count())
    st = make_data(a)
    tr = OracleTrainer(st, a.workload, bs=a.bs)
    s0 = int(st.n_events * 0.4)
    # bounded sample: the step is the first `sample_bs` events of each batch, sized from a probe step
    t0 = time.perf_counter()
    tr.train_step(s0, s0 + 256)
    per_event = (time.perf_counter() - t0) / 256
    budget = 150.0 / max(a.steps + a.warmup, 1)
    sample_bs = int(min(a.bs, max(256, budget / per_event)))
    pos = s0 + a.bs
    for _ in range(a.warmup):
        tr.train_step(pos, pos + sample_bs); pos += a.bs
    t0 = time.perf_counter()
    for _ in range(a.steps):
        tr.train_step(pos, pos + sample_bs); pos += a.bs
    dt = time.perf_counter() - t0
    v = a.steps * sample_bs / dt
    sample = f"{a.steps} steps x first {sample_bs} events of each {a.bs}-event batch (after {a.warmup} warm-up)"
    print(json.dumps({"impl": "reference", "metric": "train_events_per_sec", "value": v, "unit": "events/s",
                      "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
                      "cpu_baseline": {"value": v, "unit": "events/s", "cores": torch.get_num_threads(),
                                       "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def algorithmic_work(name, args, n_uniq):
    """(unit, amount) of algorithmic work of one C-ABI call (DESIGN.md section 'Kernels')."""
    if name in ("pfo_linear_f32", "pfo_linear_bf16"):
        M, m_dev, N, K = args[11], args[12], args[13], args[14]
        if m_dev:
            M = min(M, n_uniq)
        return "flop", 2.0 * M * N * K
    if name == "pfo_wgrad_f32":
        M, m_dev, N, K = args[5], args[6], args[7], args[8]
        if m_dev:
            M = min(M, n_uniq)
        return "flop", 2.0 * M * N * K
    if name == "pfo_attn_nbr_fwd":
        Q, n, d, F, H, ekp = args[9:15]
        return "byte", Q * (2 * H * ekp * 4 + n * (4 * d + 4 * F + 12) + H * n * 4)
    if name == "pfo_attn_nbr_bwd":
        Q, n, d, F, H, ekp = args[12:18]
        return "byte", Q * (3 * H * ekp * 4 + n * (2 * 4 * d + 4 * F + 12) + H * n * 4)
    if name == "pfo_neighbor_sample":
        Q, n = args[6], args[7]
        return "byte", Q * (16 + 8 * 10 + 28 * n)
    return None, 0.0


def profile_kernels(step_fn, n_steps, n_uniq_fn):
    """Per-entry-point device time with CUDA events on the launching stream (a separate pass)."""
    from pfotgnrec_b200 import _lib
    records = []
    orig = _lib.call

    def timed_call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *args)
        e1.record()
        records.append((name, args, e0, e1))

    _lib.call = timed_call
    try:
        for i in range(n_steps):
            step_fn(i)
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
    n_uniq = n_uniq_fn()
    agg = {}
    for name, args, e0, e1 in records:
        unit, amt = algorithmic_work(name, args, n_uniq)
        a = agg.setdefault(name, {"ms": 0.0, "calls": 0, "flop": 0.0, "byte": 0.0})
        a["ms"] += e0.elapsed_time(e1)
        a["calls"] += 1
        if unit:
            a[unit] += amt
    return agg


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    _lib.load()
    st = make_data(a)
    tc = TrainConfig(model=a.workload, bs=a.bs, gemm_mode=a.gemm)
    if world > 1:
        from pfotgnrec_b200.dist import ShardedTrainer
        tr = ShardedTrainer(st, tc, dev, rank, world)
    else:
        tr = PfoTrainer(st, tc, device=dev)
    bs = a.bs
    s0 = int(st.n_events * 0.4)                      # deep enough that every hot node has history
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    pos = [s0]

    def step(_i=None):
        s = pos[0]
        pos[0] += bs * world if world > 1 else bs
        return tr.train_step(s, s + bs) if world == 1 else tr.train_step(s, s + bs * world)

    for _ in range(a.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _lib.LAUNCHES
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(a.steps):
        flush.fill_(1)                               # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = _lib.LAUNCHES - launches0
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    events_per_step = bs * world
    value = a.steps * events_per_step / (total_ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D of the batch + D2H of the loss per step)
    e2e = None
    if world == 1:
        host = tr.make_host_batches(pos[0], a.steps + 1, bs) if hasattr(tr, "make_host_batches") else None
    else:
        host = None
    if host is not None:
        tr.train_step_host(host[0])                  # warm the path
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for hb in host[1:]:
            l = tr.train_step_host(hb)
            _ = float(l.item())                      # device -> host read of the step's result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": a.steps * bs / dt, "unit": "events/s", "h2d_bytes_per_step": int(host[1]["nbytes"]),
               "d2h_bytes_per_step": 4}

    # ---- per-kernel pass for the roofline of the dominant kernel
    roofline, kernels = None, None
    if not a.no_profile and world == 1:
        agg = profile_kernels(step, 3, lambda: int(tr.tgn.memory.state.n_unique.item()) if tr.tgn.use_memory else 0)
        tot = sum(v["ms"] for v in agg.values())
        kernels = {k: {"ms_per_step": v["ms"] / 3, "share": v["ms"] / tot, "calls_per_step": v["calls"] / 3}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        name, v = top
        if v["flop"] > 0:
            ach = v["flop"] / (v["ms"] * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": None,
                        "note": "fp32 FFMA path measured against the bf16 tensor peak (measured, sustained)"
                        if a.gemm == "fp32" else "bf16 tcgen05 path, of measured sustained peak"}
        elif v["byte"] > 0:
            ach = v["byte"] / (v["ms"] * 1e-3) / 1e9
            peak = peaks.get("hbm_gbs", 6650.0)
            roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None, "note": "of measured copy bandwidth"}

    # ---- eval users/sec (the second half of the metric): full ranking over all stocks
    eval_users = None
    if world == 1 and a.eval_steps > 0:
        ebs = a.eval_bs
        p = pos[0]
        tr.eval_step(p, p + ebs); p += ebs
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.eval_steps):
            tr.eval_step(p, p + ebs); p += ebs
        e1.record()
        torch.cuda.synchronize()
        eval_users = a.eval_steps * ebs / (e0.elapsed_time(e1) * 1e-3)

    # ---- CPU baseline (oracle port) on the host cores, bounded sample, rank 0 only
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle.train_loop import OracleTrainer
        torch.set_num_threads(os.cpu_count())
        otr = OracleTrainer(st, a.workload, bs=bs)
        sb = 1024
        otr.train_step(s0, s0 + sb)
        t0 = time.perf_counter()
        n_cpu = 0
        while time.perf_counter() - t0 < 15.0:
            otr.train_step(s0 + (n_cpu + 1) * bs, s0 + (n_cpu + 1) * bs + sb)
            n_cpu += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu * sb / dt, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n_cpu} steps x first {sb} events of a {bs}-event batch, ~15 s, after 1 warm-up step"}

    if rank == 0:
        out = {"metric": "train_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.gemm == "fp32" else "bf16",
               "data": "synthetic",
               "config": {"workload": workload_name(a), "l2": "flushed between timed steps (256 MiB write)",
                          "global_batch": events_per_step, "parallelism": f"node-sharded x{world}" if world > 1 else "1 GPU",
                          "timing": "sum of per-step CUDA-event durations, max over ranks"},
               "clocks": clk, "e2e": e2e, "gpu_launches": launches, "wall_s": wall,
               "roofline": roofline, "cpu_baseline": cpu, "eval_users_per_sec": eval_users, "kernels": kernels}
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

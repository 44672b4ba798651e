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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ours", choices=["ours", "tgn", "jodie", "dyrep", "tgat"])
    ap.add_argument("--bs", type=int, default=8192)
    ap.add_argument("--layers", type=int, default=1, help="attention layers (BASELINE config 3, TGAT: 2)")
    ap.add_argument("--neighbors", type=int, default=10, help="sampled temporal neighbours (BASELINE config 3, TGAT: 20)")
    ap.add_argument("--users", type=int, default=100000)
    ap.add_argument("--items", type=int, default=1000)
    ap.add_argument("--events", type=int, default=5000000)
    ap.add_argument("--days", type=int, default=200)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tf32", "bf16", "simt"],
                    help="fp32 = 3xTF32 on tcgen05 (1e-5 contract, default); tf32 / bf16 = 2e-2 contract; simt = FFMA")
    ap.add_argument("--dropout", type=float, default=0.1, help="attention dropout (reference main.py:29 default)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of one CUDA graph")
    ap.add_argument("--parallelism", default="sharded", choices=["replicated", "sharded"],
                    help="N > 1: node-sharded state (owner = node mod N) with device-planned all-to-all routing "
                         "(pfotgnrec_b200/dist.py, the default), or replicated state + data-parallel interactions")
    ap.add_argument("--sharded-1gpu", action="store_true",
                    help="N = 1: run the node-sharded trainer on one rank (every exchange degenerates to a copy): isolates "
                         "the cost of the routing kernels from the cost of the collectives")
    ap.add_argument("--procedural", action="store_true",
                    help="GPU-resident procedural stream (pfotgnrec_b200/synth_device.py) instead of the host-built one: the "
                         "only way to the scale configuration; runs the node-sharded trainer at any N (N = 1 included)")
    ap.add_argument("--config4", action="store_true",
                    help="BASELINE config 4: --procedural with 10^7 users x 5 000 stocks x 10^9 interactions, global batch "
                         "65 536 (= --bs 8192 on 8 GPUs)")
    ap.add_argument("--eval-steps", type=int, default=8)
    ap.add_argument("--eval-bs", type=int, default=512, help="users per evaluation batch and GPU (reference --bs default)")
    ap.add_argument("--large-bs", type=int, default=65536,
                    help="also time the step at this batch size (BASELINE config 4's) and report its per-kernel rooflines "
                         "under `large_batch`: at bs 8192 every kernel runs 10-40 us and is bound by launch / pipeline-fill "
                         "latency, not by HBM; 0 disables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true",
                    help="for ncu --profile-from-start off: after the warm-up, bracket ONE eager (kernel by kernel) training "
                         "step with cudaProfilerStart/Stop and exit; numbers printed under a profiler are not bench values")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "port"],
                    help="--impl reference: auto = the unmodified reference when its tree is staged, else the oracle port")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds for the timed + warm-up steps")
    ap.add_argument("--eval-budget", type=float, default=30.0, help="--impl reference: seconds for the evaluation sample")
    ap.add_argument("--cpu-budget", type=float, default=30.0,
                    help="seconds of reference steps for the GPU arm's cpu_baseline (run as a subprocess of --impl reference)")
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()
    if a.config4:
        a.procedural, a.users, a.items, a.events = True, 10_000_000, 5_000, 1_000_000_000
    return a


def workload_name(a):
    if a.procedural:
        return (f"{'PfoTGNRec' if a.workload == 'ours' else a.workload} train step "
                f"(d=64, {a.layers} layer{'s' if a.layers > 1 else ''}, {a.neighbors} neighbours, 2 heads"
                f"{', MV sampling K=20' if a.workload == 'ours' else ''}), "
                f"GPU-resident procedural {a.users}-user x {a.items}-stock x {a.events}-event stream, bs={a.bs} per GPU")
    return (f"{'PfoTGNRec' if a.workload == 'ours' else a.workload} train step "
            f"(d=64, {a.layers} layer{'s' if a.layers > 1 else ''}, {a.neighbors} neighbours, 2 heads"
            f"{', MV sampling K=20' if a.workload == 'ours' else ''}), "
            f"synthetic {a.users}-user x {a.items}-stock x {a.events}-event stream, bs={a.bs}")


def make_data(a):
    from pfotgnrec_b200.synth import make_stream
    return make_stream(n_users=a.users, n_items=a.items, n_events=a.events, n_days=a.days, seed=0, ts_mode="nbg")


class StreamCursor:
    """Start / end of consecutive batches inside [lo, hi) of the stream.  When the next batch would run past the end
    the cursor wraps to `lo`: a long run on 8 GPUs (global batch 65 536) consumes more events than the 5 M-event
    stream holds after the start offset, and the region is then replayed (same graph, same shapes, same work)."""

    def __init__(self, lo, hi):
        self.lo, self.hi, self.pos = int(lo), int(hi), int(lo)

    def take(self, n):
        n = int(n)
        if n > self.hi - self.lo:
            raise SystemExit(f"a batch of {n} events does not fit the stream region [{self.lo}, {self.hi})")
        if self.pos + n > self.hi:
            self.pos = self.lo
        s = self.pos
        self.pos += n
        return s, s + n


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a polling thread (2 ms period --
    the timed region is tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop, self.th, self.h, self.nv = threading.Event(), None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self._stop.set()
        self.th.join(timeout=1.0)
        reasons = sorted(k for k, bit in self.BAD.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml, 2 ms polling during the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            f = [float(x) for x in out.strip().split(",")]
            return {"sm_mhz": f[0], "sm_max_mhz": f[1], "reasons": [], "samples": 1, "source": "nvidia-smi (after)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}


def _fit_line(t_small, n_small, t_big, n_big):
    """Seconds per batch = c0 + c1 * events, from two probe steps."""
    c1 = max((t_big - t_small) / (n_big - n_small), 1e-7)
    return max(t_small - c1 * n_small, 0.0), c1


def run_reference(a):
    """The reference arm: the UNMODIFIED reference (baseline/_ref, staged by baseline/stage_reference.py; its own TGN,
    NeighborFinder, RandEdgeSampler and the text of main.py:167-394 as the step, evaluation.py:39-258 as the evaluation)
    on the host cores, kind "reference".  Only when no reference tree is present does it fall back to the oracle port
    (oracle/train_loop.py, kind "port").  Each step is a bounded sample -- the first `sample_bs` interactions of the
    config's batch -- sized from two probe steps so that the run ends within `--ref-budget` seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from stage_reference import ref_root
    torch.set_num_threads(os.cpu_count())
    st = make_data(a)
    root = ref_root()
    kind = "reference" if root is not None and a.ref_kind != "port" else "port"
    t_init = time.perf_counter()
    if kind == "reference":
        from ref_harness import ReferenceRunner
        runner = ReferenceRunner(st, a.workload, bs=a.bs, n_layers=a.layers, n_neighbors=a.neighbors,
                                 dropout=a.dropout, threads=os.cpu_count())
        one = lambda pos, n: runner.train_step(pos // n * n, n)
    else:
        from oracle.train_loop import OracleTrainer
        runner = OracleTrainer(st, a.workload, bs=a.bs, n_layers=a.layers, n_neighbors=a.neighbors)
        one = lambda pos, n: runner.train_step(pos, pos + n)
    t_init = time.perf_counter() - t_init
    s0 = int(st.n_events * 0.4)
    # positions only move forward: the reference asserts that memory is never updated to a time in the past
    # (modules/memory_updater.py:25); if the cursor ever wraps, the memory is re-initialised like at an epoch start
    probes, p = [], s0
    hi = int(st.n_events * 0.78)
    n_a = int(max(8, min(256, (hi - s0) // 20)))     # two probe sizes; tiny test streams shrink them
    n_b = 4 * n_a
    for n in (n_a, n_b):
        one(p, n)                                    # first touch at this size (allocations, lazy portfolios)
        p += n
        t0 = time.perf_counter()
        one(p, n)
        probes.append(time.perf_counter() - t0)
        p += n
    c0, c1 = _fit_line(probes[0], n_a, probes[1], n_b)
    budget = float(a.ref_budget) / max(a.steps + a.warmup, 1)
    sample_bs = int(min(a.bs, max(min(256, a.bs), (budget - c0) / c1)))
    cur = StreamCursor(p, hi)
    last = [p]

    def step():
        pos = cur.take(a.bs)[0]
        if pos < last[0] and hasattr(runner, "reset_memory"):
            runner.reset_memory()
        last[0] = pos
        one(pos, sample_bs)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    v = a.steps * sample_bs / dt
    sample = (f"{a.steps} steps x first {sample_bs} interactions of each {a.bs}-interaction batch (after {a.warmup} warm-up "
              f"steps); step = main.py:167-394 verbatim" if kind == "reference" else
              f"{a.steps} steps x first {sample_bs} events of each {a.bs}-event batch (after {a.warmup} warm-up)")
    ev = None
    if a.eval_steps > 0 and kind == "reference":
        # evaluation.py:39-258 on the full graph, all-stock ranking + metric block: a bounded number of users
        e0 = int(st.n_events * 0.9)
        ub = 16
        _, users, t1 = runner.evaluate(e0, e0 + 2 * ub + 1, ub)          # probe (builds the full-graph finder first)
        per_user = t1 / max(users, 1)
        nb = int(max(2, min(64, float(a.eval_budget) / max(per_user * ub, 1e-6))))
        _, users, t1 = runner.evaluate(e0 + 4 * ub, e0 + 4 * ub + nb * ub + 1, ub)
        ev = {"metric": "eval_users_per_sec", "value": users / t1, "unit": "users/s", "n_items": a.items,
              "cpu_baseline": {"value": users / t1, "unit": "users/s", "cores": torch.get_num_threads(), "kind": kind,
                               "sample": f"{users} users in batches of {ub} through evaluation.py:39-258 "
                                         f"(full-graph finder, ranking over all {a.items} stocks, metric block)"}}
    print(json.dumps({"impl": "reference", "metric": "train_events_per_sec", "value": v, "unit": "events/s",
                      "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": base_config(a, a.gpus),
                      "cpu_baseline": {"value": v, "unit": "events/s", "cores": torch.get_num_threads(),
                                       "kind": kind, "sample": sample,
                                       "reference_tree": root, "init_s": t_init,
                                       "threads_note": "the reference's Python loops are single-threaded; torch / BLAS ops "
                                                       "use all cores"},
                      "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "eval": ev}))


def cpu_baseline_subprocess(a):
    """cpu_baseline of the GPU arm: the reference arm above in its OWN process (the drop-in overlay and the reference
    share module paths, so they cannot live in one interpreter), with a short budget.  Returns the parsed JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "1",
           "--ref-budget", str(a.cpu_budget), "--eval-budget", str(a.cpu_budget / 2), "--workload", a.workload,
           "--bs", str(a.bs), "--layers", str(a.layers), "--neighbors", str(a.neighbors), "--users", str(a.users),
           "--items", str(a.items), "--events", str(a.events), "--days", str(a.days), "--dropout", str(a.dropout),
           "--eval-steps", str(a.eval_steps)]
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = [l for l in p.stdout.strip().split("\n") if l.startswith("{")][-1]
        return json.loads(line)
    except Exception as exc:                         # the baseline is a reported figure: never fail the GPU line on it
        return {"error": f"{type(exc).__name__}: {exc}"}


def base_config(a, world):
    """The part of `config` both arms share: what is computed, not how."""
    return {"workload": workload_name(a), "global_batch": a.bs * max(int(world), 1), "dropout": a.dropout,
            "eval": f"full ranking over all {a.items} stocks + metric block (Recall/NDCG, delta-return/delta-Sharpe @1,3,5)"}


def algorithmic_work(name, args, n_uniq, extra):
    """(unit, amount) of ALGORITHMIC work of one C-ABI call (DESIGN.md section 4 / SURVEY.md section 8d)."""
    if name in ("pfo_linear_f32", "pfo_linear_bf16", "pfo_linear_tf32"):
        M, m_dev, N, K = args[11], args[12], args[13], args[14]
        if m_dev:
            M = min(M, n_uniq)
        return "gemm", (2.0 * M * N * K, 4.0 * M * (K + N) + 4.0 * N * K)      # (flop, compulsory bytes: A in, C out, W)
    if name in ("pfo_wgrad_f32", "pfo_wgrad_tf32"):
        M, m_dev, N, K = args[5], args[6], args[7], args[8]
        if m_dev:
            M = min(M, n_uniq)
        return "gemm", (2.0 * M * N * K, 4.0 * M * (K + N) + 4.0 * N * K)
    if name == "pfo_attn_nbr_fwd":
        Q, n, d, F, H, ekp = args[9:15]
        return "byte", Q * (2 * H * ekp * 4 + n * (4 * d + 4 * F + 12) + H * n * 4)
    if name == "pfo_attn_nbr_fwd_rows":              # query operand shared per node: its rows are a cache-resident table,
        Q, n, d, F, H, ekp = args[10:16]             # so only the row index counts as per-query traffic
        return "byte", Q * (H * ekp * 4 + 4 + n * (4 * d + 4 * F + 12) + H * n * 4)
    if name == "pfo_attn_nbr_bwd":
        Q, n, d, F, H, ekp = args[13:19]
        return "byte", Q * (3 * H * ekp * 4 + n * (2 * 4 * d + 4 * F + 12) + H * n * 4)
    if name == "pfo_neighbor_sample":
        Q, n = args[6], args[7]
        return "byte", Q * (16 + 8 * extra["log2deg"] + 28 * n)
    if name == "pfo_mv_select":
        n_ret, B, K = args[9], args[10], args[11]
        return "byte", B * ((K + 1 + extra["mean_portfolio"]) * n_ret * 8 + 4 * extra["mean_portfolio"] + 16)
    if name == "pfo_bpr":
        B, k, d = args[3], args[4], args[5]
        return "byte", B * (2 + k) * 4 * d * 2
    if name == "pfo_store_messages":
        B, d, F = args[4], args[5], args[6]
        return "byte", 2 * B * ((2 * 4 * d + 4 * F + 12) + (4 * (3 * d + F) + 5))
    if name == "pfo_gather_state":                   # rows of the unique nodes: memory + pending message, read and written
        u_max, d, raw = args[2], args[3], args[4]
        return "byte", min(u_max, n_uniq) * (2 * 4 * (d + raw) + 4 + 4 + 1 + 12)
    if name in ("pfo_cell_forward", "pfo_cell_backward"):
        u_max, d, cell = args[2], args[3], args[4]
        g = {0: 3, 1: 1}.get(cell, 0) * d
        return "byte", min(u_max, n_uniq) * 4 * (2 * g + (4 * d if name == "pfo_cell_forward" else d + 2 * g + d))
    if name == "pfo_time_embedding_fwd":
        Q, d = args[2], args[4]
        return "byte", Q * (12 + 4 + 2 * 4 * d + 4)
    if name == "pfo_time_embedding_bwd":
        Q, d = args[1], args[2]
        return "byte", Q * (4 + 4 + 3 * 4 * d)
    return None, 0.0


def profile_kernels(step_fn, n_steps, n_uniq_fn, extra):
    """Per-entry-point device time with CUDA events on the launching stream, in a separate eager pass (the timed
    region replays a CUDA graph, which has no per-kernel hooks).  A device-side sleep is queued ahead of every
    profiled step so the host runs ahead of the GPU and the events bracket kernels, not launch gaps."""
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
            torch.cuda._sleep(40_000_000)            # ~20 ms at 1.9 GHz
            step_fn(i)
            torch.cuda.synchronize()
    finally:
        _lib.call = orig
    n_uniq = n_uniq_fn()
    agg = {}
    for name, args, e0, e1 in records:
        unit, amt = algorithmic_work(name, args, n_uniq, extra)
        a = agg.setdefault(name, {"ms": 0.0, "calls": 0, "flop": 0.0, "byte": 0.0})
        a["ms"] += e0.elapsed_time(e1)
        a["calls"] += 1
        if unit == "gemm":
            a["flop"] += amt[0]
            a["byte"] += amt[1]
        elif unit:
            a[unit] += amt
    return agg


def roofline_of(name, v, peaks, gemm, ncu):
    """roofline object of one entry point from its aggregated (ms, algorithmic work)."""
    if v["ms"] <= 0:
        return None
    traffic = (ncu.get(name) or {}).get("dram_bytes_per_launch")
    tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
    hpeak = peaks.get("hbm_gbs", 6650.0)
    if v["flop"] > 0 and v["flop"] / max(v["byte"], 1.0) >= tpeak * 1e12 / (hpeak * 1e9):
        ach = v["flop"] / (v["ms"] * 1e-3) / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                "traffic": traffic, "launches": v["calls"],
                "note": "peak = measured sustained dense bf16 (MEASURED_PEAKS.json)"}
    if v["flop"] > 0:
        # the tall-skinny contractions of this path (K, N <= 328, fp32 operands in HBM) sit at ~25-50 FLOP/byte, far
        # left of the ~210 FLOP/byte ridge: they are bound by moving the activations, not by the tensor pipe
        ach = v["byte"] / (v["ms"] * 1e-3) / 1e9
        tf = v["flop"] / (v["ms"] * 1e-3) / 1e12
        mode = {"fp32": "3xTF32 on tcgen05 (3 MMAs per algorithmic MAC)", "tf32": "TF32 on tcgen05",
                "bf16": "bf16 on tcgen05", "simt": "fp32 FFMA"}[gemm]
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hpeak, "unit": "GB/s", "frac": ach / hpeak,
                "traffic": traffic, "launches": v["calls"], "flop_per_byte": v["flop"] / v["byte"],
                "tensor_tflops": tf, "tensor_frac": tf / tpeak,
                "note": mode + "; GEMM below the roofline ridge, bytes = A in + C out + W per launch; peak = measured "
                        "copy bandwidth (MEASURED_PEAKS.json); traffic = mean DRAM bytes per launch from the committed "
                        "ncu --set full capture"}
    if v["byte"] > 0:
        ach = v["byte"] / (v["ms"] * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hpeak, "unit": "GB/s", "frac": ach / hpeak,
                "traffic": traffic, "launches": v["calls"], "note": "peak = measured copy bandwidth (MEASURED_PEAKS.json)"}
    return None


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    if os.environ.get("PFO_HANG_DUMP_S"):            # debugging aid for multi-rank runs: where every thread sits, then exit
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PFO_HANG_DUMP_S"]), exit=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    from pfotgnrec_b200 import _lib
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    _lib.load()
    one_rank = world == 1 and (a.procedural or a.sharded_1gpu)
    if one_rank:                                     # the sharded trainer's exchange needs a process group, even of one
        import torch.distributed as dist
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29700 + os.getpid() % 200}", rank=0,
                                world_size=1, device_id=dev)
    if a.procedural:
        from pfotgnrec_b200.synth_device import DeviceStream
        st = DeviceStream(a.users, a.items, a.events, n_days=a.days, seed=0, device=dev)
    else:
        st = make_data(a)
    tc = TrainConfig(model=a.workload, bs=a.bs, gemm_mode=a.gemm, dropout=a.dropout, cuda_graph=not a.no_graph,
                     n_layers=a.layers, n_neighbors=a.neighbors)
    if one_rank or (world > 1 and a.parallelism == "sharded"):
        from pfotgnrec_b200.dist import ShardedTrainer
        tr = ShardedTrainer(st, tc, dev, rank, world)
        if a.no_graph:
            tr.tc.cuda_graph = False
    elif world > 1:
        from pfotgnrec_b200.trainer import ReplicatedTrainer
        tr = ReplicatedTrainer(st, tc, dev, rank, world)
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

    cur = StreamCursor(s0, st.n_events)
    if a.procedural:                                 # evaluate the batches' columns ahead of the timed loops
        pre = StreamCursor(s0, st.n_events)
        tr.prefetch([pre.take(bs * world) for _ in range(a.warmup + 2 * a.steps + 8)])

    def step(_i=None):
        s, e = cur.take(bs * world)                  # the GLOBAL batch; each rank embeds its slice of it
        return tr.train_step(s, e)

    for _ in range(a.warmup):
        step()
    barrier()
    if a.ncu_step:
        tr.tc.cuda_graph = False
        step()                                       # one un-profiled eager step (lazy allocations of the eager path)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_step": True, "note": "one eager training step was bracketed for the profiler"}))
        return
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
    if hasattr(tr, "ex"):
        tr.ex.check_overflow()                       # a calibrated bucket that overflowed would have dropped rows
    events_per_step = bs * world
    value = a.steps * events_per_step / (total_ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D of the batch + D2H of the loss per step)
    e2e = None
    host = None
    if hasattr(tr, "make_host_batches"):             # one cursor position per batch (the cursor may wrap in between)
        host = [tr.make_host_batches(cur.take(bs * world)[0], 1, bs)[0] for _ in range(a.steps + 1)]
    if host is not None:
        tr.train_step_host(host[0])                  # warm the path
        barrier()
        dt = 0.0
        for hb in host[1:]:
            flush.fill_(1)                           # same L2 flush as the device-timed loop, outside the step's clock
            torch.cuda.synchronize()
            t0 = time.perf_counter()                 # host clock around the call a user makes: H2D of the batch ->
            l = tr.train_step_host(hb)               # step -> D2H of the loss (the read synchronises the device)
            _ = float(l.item())
            dt += time.perf_counter() - t0
        barrier()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": a.steps * events_per_step / dt, "unit": "events/s",
               "h2d_bytes_per_step": int(host[1]["nbytes"]) * world, "d2h_bytes_per_step": 4 * world,
               "timing": "sum of per-step host wall-clock durations (pinned host batch -> one H2D copy -> step -> loss "
                         "read back), L2 flushed between steps outside the clock, max over ranks"}

    # ---- per-kernel pass (eager launches, CUDA events per C-ABI call) for the rooflines
    roofline, kernels, rooflines = None, None, None
    if not a.no_profile and world == 1 and not a.procedural:
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except (OSError, ValueError):
            pass
        deg = np.diff(tr.csr_train.rowptr.cpu().numpy())
        hot = np.concatenate([st.sources[s0:s0 + 4 * bs], st.destinations[s0:s0 + 4 * bs]])
        extra = {"log2deg": float(np.mean(np.ceil(np.log2(deg[hot] + 1.0)))),
                 "mean_portfolio": float(np.mean(np.diff(st.port_ptr[s0:s0 + 4 * bs + 1])))}
        graph_mode, tr.tc.cuda_graph = tr.tc.cuda_graph, False
        try:
            agg = profile_kernels(step, 3, lambda: int(tr.tgn.memory.state.n_unique.item()) if tr.tgn.use_memory else 0,
                                  extra)
        finally:
            tr.tc.cuda_graph = graph_mode
        tot = sum(v["ms"] for v in agg.values())
        kernels = {k: {"ms_per_step": v["ms"] / 3, "share": v["ms"] / tot, "calls_per_step": v["calls"] / 3}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        rooflines = {}
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            r = roofline_of(k, v, peaks, a.gemm, ncu)
            if r is not None:
                r["share_of_step"] = v["ms"] / tot
                rooflines[k] = r
        roofline = next(iter(rooflines.values()), None)      # the most expensive entry point with a roofline model

    # ---- the same step at the scale configuration's batch size: what the kernels reach once a launch carries enough
    # rows to leave the latency regime (extra information; `value` above stays the bs-8192 headline)
    large = None
    if (a.large_bs > 0 and world == 1 and rooflines is not None and a.large_bs != bs
            and 2 * a.large_bs < st.n_events - s0):
        BL = a.large_bs

        def lstep(_i=None):
            s, e = cur.take(BL)
            return tr.train_step(s, e)

        for _ in range(4):                           # two eager steps, capture, one replay
            lstep()
        torch.cuda.synchronize()
        levs = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lstep(); e1.record()
            levs.append((e0, e1))
        torch.cuda.synchronize()
        lms = sum(x.elapsed_time(y) for x, y in levs) / len(levs)
        graph_mode, tr.tc.cuda_graph = tr.tc.cuda_graph, False
        try:
            lagg = profile_kernels(lstep, 2, lambda: int(tr.tgn.memory.state.n_unique.item()) if tr.tgn.use_memory else 0, extra)
        finally:
            tr.tc.cuda_graph = graph_mode
        ltot = sum(v["ms"] for v in lagg.values())
        lroof = {}
        for k, v in sorted(lagg.items(), key=lambda kv: -kv[1]["ms"]):
            r = roofline_of(k, v, peaks, a.gemm, {})
            if r is not None:
                lroof[k] = {"bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"],
                            "frac": r["frac"], "share_of_step": v["ms"] / ltot, "ms_per_step": v["ms"] / 2}
        large = {"global_batch": BL, "value": BL / (lms * 1e-3), "unit": "events/s", "ms_per_step": lms, "steps": 5,
                 "rooflines": lroof,
                 "note": "same model, stream and timing rules at BASELINE config 4's batch size (graph replay, L2 flushed)"}

    # ---- eval users/sec (the second half of the metric): full ranking over all stocks, users split over the ranks
    eval_users, eval_roofline = None, None
    if a.eval_steps > 0 and hasattr(tr, "eval_step"):
        ebs = a.eval_bs * world                      # global evaluation batch
        for _ in range(3):                           # two eager steps, then the graph is captured
            tr.eval_step(*cur.take(ebs))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.eval_steps):
            tr.eval_step(*cur.take(ebs))
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        eval_users = a.eval_steps * ebs / (ms * 1e-3)
        if world == 1 and not a.no_profile and rooflines is not None:
            # per-kernel rooflines of the evaluation step (Q = users x (2 + N_ITEMS) queries), eager pass like the training one
            graph_mode, tr.tc.cuda_graph = tr.tc.cuda_graph, False
            try:
                eagg = profile_kernels(lambda _i: tr.eval_step(*cur.take(ebs)), 2,
                                       lambda: int(tr.tgn.memory.state.n_unique.item()) if tr.tgn.use_memory else 0, extra)
            finally:
                tr.tc.cuda_graph = graph_mode
            etot = sum(v["ms"] for v in eagg.values())
            for k, v in sorted(eagg.items(), key=lambda kv: -kv[1]["ms"]):
                r = roofline_of(k, v, peaks, a.gemm, {})
                if r is not None:
                    r["share_of_step"] = v["ms"] / etot
                    r["ms_per_step"] = v["ms"] / 2
                    eval_roofline = r
                    break

    # ---- CPU baseline: the unmodified reference on the host cores (its own process), bounded sample, rank 0 only
    cpu, eval_cpu = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and not a.procedural:
        ref = cpu_baseline_subprocess(a)
        cpu = ref.get("cpu_baseline") or ref
        eval_cpu = (ref.get("eval") or {}).get("cpu_baseline")

    if rank == 0:
        ran_graph = bool(getattr(tr, "_graph_ok", lambda _b: False)(bs))
        cfg = base_config(a, world)
        cfg.update({"l2": "flushed between timed steps (256 MiB write)",
                    "parallelism": ("1 GPU" if world == 1 and not one_rank else
                                    (f"node-sharded x{world} (owner = node mod {world}, device-planned all-to-all)"
                                     if a.parallelism == "sharded" or one_rank
                                     else f"replicated state, data-parallel x{world}")),
                    "timing": "sum of per-step CUDA-event durations, max over ranks",
                    "gemm_mode": a.gemm, "cuda_graph": ran_graph})
        if hasattr(tr, "ex"):
            cfg["exchange_transport"] = ("peer-memory stores + flag barrier (CUDA IPC arenas over NVLink)"
                                         if tr.ex.transport == "peer" else "NCCL all_to_all_single")
        ev_block = None
        if eval_users is not None:
            ev_block = {"metric": "eval_users_per_sec", "value": eval_users, "unit": "users/s", "n_items": a.items,
                        "users_per_batch": a.eval_bs * world, "steps": a.eval_steps, "roofline": eval_roofline,
                        "cpu_baseline": eval_cpu,
                        "note": "full ranking over all stocks per user (evaluation.py:84-138) + metric block (:127-258), "
                                "graph-replayed, CUDA events, max over ranks"}
        out = {"metric": "train_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.gemm == "fp32" else "bf16",
               "data": "synthetic", "config": cfg,
               "clocks": clk, "e2e": e2e, "gpu_launches": launches, "wall_s": wall,
               "roofline": roofline, "cpu_baseline": cpu, "eval": ev_block, "eval_users_per_sec": eval_users,
               "kernels": kernels, "rooflines": rooflines, "large_batch": large}
        print(json.dumps(out))
    if world > 1 or one_rank:
        # the captured graphs of the sharded mode hold NCCL kernels; tearing the communicator down under them can block
        # for minutes -- everything is printed, so synchronise, meet the other ranks and leave without the teardown
        sys.stdout.flush()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()

"""Where does a training step's time go?  Device time per step vs host time per step, with and
without the L2 flush, on the bench workload (debug aid; prints to stdout)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfotgnrec_b200.synth import make_stream
from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig

ev = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
st = make_stream(100000, 1000, ev, 200, seed=0, ts_mode="nbg")
tr = PfoTrainer(st, TrainConfig(model="ours", bs=bs))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pos = int(ev * 0.4)
for _ in range(3):
    tr.train_step(pos, pos + bs); pos += bs
torch.cuda.synchronize()
for mode in ("noflush_nosync", "flush_nosync", "noflush_sync", "flush_sync"):
    evs, t0 = [], time.perf_counter()
    for _ in range(10):
        if mode.startswith("flush"):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr.train_step(pos, pos + bs); e1.record(); pos += bs
        if mode.endswith("_sync"):
            torch.cuda.synchronize()
        evs.append((e0, e1))
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(mode, "host enqueue ms/step %.2f  wall ms/step %.2f  device ms/step %s" % (
        1e3 * t_enq / 10, 1e3 * t_all / 10, " ".join("%.2f" % a.elapsed_time(b) for a, b in evs)))
if "--ncu" in sys.argv:
    torch.cuda.profiler.start()
    tr.train_step(pos, pos + bs)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

"""torchrun --nproc-per-node G tools/check_sharded.py [model] : node-sharded path == 1-GPU path on the same global
batches -- loss, parameter gradients, memory, last_update, pending-message flags per training step (two eager steps,
the capture, graph replays, a ragged tail batch), then evaluation ranks / scores -- and no bucket overflow."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pfotgnrec_b200.synth import make_stream
from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
from pfotgnrec_b200.dist import ShardedTrainer


def trainable(tr):
    return [(k, p) for k, p in tr.tgn.named_parameters() if p.requires_grad]


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("PFO_HANG_DUMP_S", "100")), exit=True)   # where a hang sits
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    model = sys.argv[1] if len(sys.argv) > 1 else "ours"
    graph = "--no-graph" not in sys.argv
    layers, nbrs = (2, 5) if model == "tgat" else (1, 10)
    st = make_stream(n_users=3000, n_items=120, n_events=40000, n_days=40, seed=2, ts_mode="small")
    bs = 256
    kw = dict(model=model, lr=0.0, n_layers=layers, n_neighbors=nbrs, cuda_graph=graph)   # lr 0: weights stay equal
    sh = ShardedTrainer(st, TrainConfig(bs=bs, **kw), dev, rank, world)
    single = PfoTrainer(st, TrainConfig(bs=bs * world, **kw), device=dev)
    for (k, a), (_, b) in zip(trainable(sh), trainable(single)):
        assert torch.equal(a, b), k
    s0, worst = 12000, 0.0
    spans = [(s0 + i * bs * world, s0 + (i + 1) * bs * world) for i in range(6)]
    spans.append((spans[-1][1], spans[-1][1] + bs * world - 3))          # ragged tail: slices differ by one interaction
    for i, (s, e) in enumerate(spans):
        l_loc = sh.train_step(s, e).clone()
        ls, le = sh._slice(s, e)
        l_sh = l_loc * (le - ls) / float(e - s)                           # mean over the global batch
        dist.all_reduce(l_sh)
        l_sh = float(l_sh.item())
        l_1 = float(single.train_step(s, e).item())
        assert abs(l_sh - l_1) < 1e-5 * max(1.0, abs(l_1)), (i, l_sh, l_1)
        if rank == 0:
            print(f"step {i}: loss {l_sh:.6f} == {l_1:.6f}", flush=True)
        for (k, a), (_, b) in zip(trainable(sh), trainable(single)):
            ga = a.grad if a.grad is not None else torch.zeros_like(a)
            gb = b.grad if b.grad is not None else torch.zeros_like(b)
            scale = max(float(gb.abs().max()), 1e-3)
            err = float((ga - gb).abs().max()) / scale
            worst = max(worst, err)
            assert err < 1e-4, (i, k, err)
        merr = 0.0
        if sh.tgn.use_memory:
            mem, lu, pv = sh.gather_memory()
            s1 = single.tgn.memory.state
            merr = float((mem - s1.memory).abs().max() / s1.memory.abs().max().clamp(min=1e-30))
            assert merr < 1e-5, (i, merr)
            assert torch.equal(lu, s1.last_update), i
            assert torch.equal(pv.to(torch.uint8), s1.pend_valid), i
    # evaluation: this rank's users against all stocks == the matching rows of the 1-GPU evaluation batch
    ebs, everr = 64, 0.0
    e0 = spans[-1][1]
    for i in range(4):
        s, e = e0 + i * ebs * world, e0 + (i + 1) * ebs * world
        r_sh = sh.eval_step(s, e)
        r_1 = single.eval_step(s, e)
        ls, le = sh._slice(s, e)
        assert torch.equal(r_sh[2], r_1[2][ls - s:le - s]), i             # candidates (Philox keyed by global position)
        sc = r_1[3][ls - s:le - s]
        everr = max(everr, float((r_sh[3] - sc).abs().max() / sc.abs().max().clamp(min=1e-30)))
        assert everr < 1e-5, (i, everr)
        assert float((r_sh[0] != r_1[0][ls - s:le - s]).float().mean()) < 0.02, i   # ranks (near-ties may flip)
    if sh.metrics is not None:
        a, b = sh.eval_summary("val"), single.eval_summary("val")
        for k, v in b.items():
            assert abs(a[k] - v) <= 2e-2 * max(1e-3, abs(v)) + 1e-6, (k, a[k], v)
    sh.ex.check_overflow()
    # timing of the replayed sharded step vs the single-GPU step on the same global batch (informative)
    torch.cuda.synchronize(); dist.barrier()
    t = {}
    for name, tr in (("sharded", sh), ("single", single)):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for i in range(20):
            s = s0 + (i % 6) * bs * world
            tr.train_step(s, s + bs * world)
        torch.cuda.synchronize()
        t[name] = (time.perf_counter() - t0) / 20 * 1e3
    if rank == 0:
        caps = {str(k): v for k, v in sh.ex.frozen.items()}
        print(f"transport {sh.ex.transport}; ", end="")
        print(f"sharded x{world} == single GPU ({model}, graph={graph}): {len(spans)} train steps (last ragged) + 4 eval "
              f"steps, loss {l_sh:.6f} vs {l_1:.6f}, worst grad rel.err {worst:.2e}, memory rel.err {merr:.2e}, "
              f"eval score rel.err {everr:.2e}; {t['sharded']:.3f} ms/step sharded vs {t['single']:.3f} ms single "
              f"(global batch {bs * world}); frozen capacities {caps}", flush=True)
    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    # captured graphs hold NCCL kernels: release them before the communicator goes away, and do not wait on teardown
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()

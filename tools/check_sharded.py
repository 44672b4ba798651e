"""torchrun --nproc-per-node G tools/check_sharded.py : node-sharded path == 1-GPU path on the same
global batches (loss, parameter gradients, memory, last_update, pending-message flags)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pfotgnrec_b200.synth import make_stream
from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
from pfotgnrec_b200.dist import ShardedTrainer


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    model = sys.argv[1] if len(sys.argv) > 1 else "ours"
    st = make_stream(n_users=3000, n_items=120, n_events=40000, n_days=40, seed=2, ts_mode="small")
    bs = 256
    tc = TrainConfig(model=model, bs=bs, lr=0.0)           # lr 0: weights stay equal, compare per-step quantities
    sh = ShardedTrainer(st, tc, dev, rank, world)
    single = PfoTrainer(st, TrainConfig(model=model, bs=bs * world, lr=0.0), device=dev)
    for (k, a), (_, b) in zip(sh.tgn.named_parameters(), single.tgn.named_parameters()):
        assert torch.equal(a, b), k
    s0, worst = 12000, 0.0
    for i in range(6):
        s, e = s0 + i * bs * world, s0 + (i + 1) * bs * world
        l_sh = sh.train_step(s, e).clone()
        dist.all_reduce(l_sh)
        l_sh = float(l_sh.item()) / world
        l_1 = float(single.train_step(s, e).item())
        assert abs(l_sh - l_1) < 1e-5 * max(1.0, abs(l_1)), (i, l_sh, l_1)
        for (k, a), (_, b) in zip(sh.tgn.named_parameters(), single.tgn.named_parameters()):
            ga = a.grad if a.grad is not None else torch.zeros_like(a)
            gb = b.grad if b.grad is not None else torch.zeros_like(b)
            scale = max(float(gb.abs().max()), 1e-3)
            err = float((ga - gb).abs().max()) / scale
            worst = max(worst, err)
            assert err < 1e-4, (i, k, err)
        mem, lu, pv = sh.gather_memory()
        s1 = single.tgn.memory.state
        merr = float((mem - s1.memory).abs().max() / s1.memory.abs().max().clamp(min=1e-30))
        assert merr < 1e-5, (i, merr)
        assert torch.equal(lu, s1.last_update), i
        assert torch.equal(pv.to(torch.uint8), s1.pend_valid), i
    if rank == 0:
        print(f"sharded x{world} == single GPU ({model}): 6 steps, loss {l_sh:.6f} vs {l_1:.6f}, "
              f"worst grad rel.err {worst:.2e}, memory rel.err {merr:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""torchrun --nproc-per-node G tools/check_replicated.py [model] : replicated-state data-parallel path == 1-GPU path
on the same global batches (loss, parameter gradients, memory, last_update, pending messages), eager and graphed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pfotgnrec_b200.synth import make_stream
from pfotgnrec_b200.trainer import PfoTrainer, ReplicatedTrainer, TrainConfig


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    model = sys.argv[1] if len(sys.argv) > 1 else "ours"
    st = make_stream(n_users=3000, n_items=120, n_events=40000, n_days=40, seed=2, ts_mode="small")
    bs = 256
    for graph in (False, True):
        rp = ReplicatedTrainer(st, TrainConfig(model=model, bs=bs, lr=0.0, cuda_graph=graph), dev, rank, world)
        single = PfoTrainer(st, TrainConfig(model=model, bs=bs * world, lr=0.0, cuda_graph=False), device=dev)
        s0, worst = 12000, 0.0
        for i in range(6):
            s, e = s0 + i * bs * world, s0 + (i + 1) * bs * world
            l_rp = rp.train_step(s, e).clone()
            dist.all_reduce(l_rp)
            l_rp = float(l_rp.item()) / world
            l_1 = float(single.train_step(s, e).item())
            assert abs(l_rp - l_1) < 1e-5 * max(1.0, abs(l_1)), (graph, i, l_rp, l_1)
            for (k, a), (_, b) in zip(rp.tgn.named_parameters(), single.tgn.named_parameters()):
                if not a.requires_grad:
                    continue
                gb = b.grad if b.grad is not None else torch.zeros_like(b)
                scale = max(float(gb.abs().max()), 1e-3)
                err = float((a.grad - gb).abs().max()) / scale
                worst = max(worst, err)
                assert err < 1e-4, (graph, i, k, err)
            sa, sb = rp.tgn.memory.state, single.tgn.memory.state
            merr = float((sa.memory - sb.memory).abs().max() / sb.memory.abs().max().clamp(min=1e-30))
            assert merr < 1e-5, (graph, i, merr)
            assert torch.equal(sa.last_update, sb.last_update) and torch.equal(sa.pend_valid, sb.pend_valid), (graph, i)
            assert torch.equal(sa.pend_ts, sb.pend_ts), (graph, i)
        if graph:
            assert rp._graphs[bs].graph is not None
        if rank == 0:
            print(f"replicated x{world} ({'graph' if graph else 'eager'}) == single GPU ({model}): 6 steps, loss "
                  f"{l_rp:.6f} vs {l_1:.6f}, worst grad rel.err {worst:.2e}, memory rel.err {merr:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

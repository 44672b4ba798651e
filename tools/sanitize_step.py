"""compute-sanitizer target: every kernel family of the path on small shapes, launched kernel by kernel (no CUDA graph).

    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py

Three PfoTGNRec training steps + one evaluation step (K5, K1, compaction, K3 linear / cell, fold, K4 fwd / bwd, wgrad,
K6, K2, eval score / metrics), one step each of the RNN + time-embedding and the 2-layer uniform-neighbour models, and the
routing kernels of the node-sharded mode on one rank.  Prints the losses; the sanitizer's own summary is the result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pfotgnrec_b200.synth import make_stream
from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig

st = make_stream(n_users=400, n_items=60, n_events=4000, n_days=20, seed=1, ts_mode="nbg")
out = {}
for model, layers, nbrs in (("ours", 1, 10), ("jodie", 1, 10), ("tgat", 2, 5)):
    tr = PfoTrainer(st, TrainConfig(model=model, bs=96, n_layers=layers, n_neighbors=nbrs, dropout=0.1, cuda_graph=False),
                    device="cuda:0")
    out[model] = [float(tr.train_step(1500 + 96 * i, 1596 + 96 * i).item()) for i in range(3 if model == "ours" else 1)]
    if model == "ours":
        tr.eval_step(3000, 3032)
    torch.cuda.synchronize()
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29900 + os.getpid() % 90}", rank=0, world_size=1,
                        device_id=torch.device("cuda", 0))
from pfotgnrec_b200.dist import ShardedTrainer
sh = ShardedTrainer(st, TrainConfig(model="ours", bs=96, dropout=0.1, cuda_graph=False), "cuda:0", 0, 1)
out["sharded_1rank"] = [float(sh.train_step(1500 + 96 * i, 1596 + 96 * i).item()) for i in range(2)]
sh.eval_step(3000, 3032)
sh.ex.freeze()          # calibrated capacities: the next steps go through the peer-memory transport (push + barrier)
out["sharded_1rank_peer"] = [float(sh.train_step(1700 + 96 * i, 1796 + 96 * i).item()) for i in range(2)]
sh.eval_step(3032, 3064)
out["transport"] = sh.ex.transport
torch.cuda.synchronize()
sh.ex.check_overflow()
print("sanitize_step:", out, flush=True)
os._exit(0)

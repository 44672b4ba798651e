"""Node-sharded multi-GPU path on 2 GPUs: tools/check_sharded.py under torchrun (needs >= 2 GPUs; the single-GPU
driver run skips it, `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("model", ["ours", "tgn", "jodie", "tgat"])
def test_sharded_equals_single_gpu_on_two_ranks(model):
    port = 29600 + os.getpid() % 300
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "check_sharded.py"), model],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert "== single GPU" in p.stdout, tail
    print(p.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def one_rank_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29800 + os.getpid() % 150}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def _run_steps(tr, spans, eval_spans):
    out = []
    for s, e in spans:
        loss = float(tr.train_step(s, e).item())
        grads = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).clone()
                           for p in tr.tgn.parameters() if p.requires_grad])
        out.append((loss, grads))
    ev = [tuple(t.clone() if t is not None else None for t in tr.eval_step(s, e)) for s, e in eval_spans]
    return out, ev


@pytest.mark.parametrize("model", ["ours", "tgat"])
def test_sharded_engine_with_one_rank_equals_the_plain_trainer(one_rank_group, model):
    """The whole node-sharded machinery on ONE rank (owner = node mod 1): device plans, slot buckets, request / reply
    rows, routed messages, calibrated capacities, graph capture -- every exchange degenerates to a copy, so the step
    must reproduce the plain single-GPU trainer on the same batches (runs in the 1-GPU driver test tier)."""
    sys.path.insert(0, ROOT)
    from pfotgnrec_b200.synth import make_stream
    from pfotgnrec_b200.trainer import PfoTrainer, TrainConfig
    from pfotgnrec_b200.dist import ShardedTrainer
    st = make_stream(n_users=1500, n_items=80, n_events=20000, n_days=30, seed=5, ts_mode="small")
    layers, nbrs = (2, 5) if model == "tgat" else (1, 10)
    tc = dict(model=model, bs=256, lr=0.0, n_layers=layers, n_neighbors=nbrs)
    sh = ShardedTrainer(st, TrainConfig(**tc), "cuda:0", 0, 1)
    pl = PfoTrainer(st, TrainConfig(**tc), device="cuda:0")
    spans = [(6000 + 256 * i, 6256 + 256 * i) for i in range(5)] + [(7280, 7280 + 251)]
    evs = [(9000 + 64 * i, 9064 + 64 * i) for i in range(4)]
    a, ea = _run_steps(sh, spans, evs)
    b, eb = _run_steps(pl, spans, evs)
    for i, ((la, ga), (lb, gb)) in enumerate(zip(a, b)):
        assert abs(la - lb) < 1e-6 * max(1.0, abs(lb)), (i, la, lb)
        assert float((ga - gb).abs().max()) <= 2e-5 * max(float(gb.abs().max()), 1e-3), i
    if sh.tgn.use_memory:
        assert torch.allclose(sh.tgn.memory.memory, pl.tgn.memory.memory, rtol=1e-5, atol=1e-7)
        assert torch.equal(sh.tgn.memory.last_update, pl.tgn.memory.last_update)
        assert torch.equal(sh.tgn.memory.state.pend_valid, pl.tgn.memory.state.pend_valid)
    for ra, rb in zip(ea, eb):
        assert torch.equal(ra[2], rb[2])                                   # candidates
        assert torch.allclose(ra[3], rb[3], rtol=1e-5, atol=1e-6)          # scores
    sh.ex.check_overflow()
    assert sh.ex.frozen                                                    # capacities were calibrated and frozen


def test_procedural_stream_trains_like_its_materialised_copy(one_rank_group):
    """The GPU-resident procedural stream of the scale configuration (synth_device.DeviceStream: columns evaluated
    from the interaction index, chunked device CSR build, feature tables drawn on the device) against the same
    interactions materialised into a host `Stream` and fed through the ordinary path."""
    sys.path.insert(0, ROOT)
    from pfotgnrec_b200.synth_device import DeviceStream
    from pfotgnrec_b200.trainer import TrainConfig
    from pfotgnrec_b200.dist import ShardedTrainer
    ds = DeviceStream(n_users=4000, n_items=150, n_events=30000, n_days=25, seed=2, device="cuda:0")
    tc = dict(model="ours", bs=256, lr=0.0)
    a = ShardedTrainer(ds, TrainConfig(**tc), "cuda:0", 0, 1)
    b = ShardedTrainer(ds.materialise(), TrainConfig(**tc), "cuda:0", 0, 1)
    assert a.n_train == b.n_train and a.split_ranges() == b.split_ranges()
    assert np.array_equal(a.universe_items, b.universe_items)
    for k in ("rowptr", "nbr", "eidx", "ts"):
        assert torch.equal(getattr(a.csr_train, k), getattr(b.csr_train, k)), k
        assert torch.equal(getattr(a.csr_full, k), getattr(b.csr_full, k)), k
    assert torch.allclose(a.tgn.edge_raw_features, b.tgn.edge_raw_features, rtol=1e-5, atol=1e-6)
    b.tgn.node_raw_features.copy_(a.tgn.node_raw_features)                 # the host path draws them from numpy
    b.tgn.edge_raw_features.copy_(a.tgn.edge_raw_features)
    spans = [(10000 + 256 * i, 10256 + 256 * i) for i in range(5)]
    evs = [(26000 + 64 * i, 26064 + 64 * i) for i in range(3)]
    ra, ea = _run_steps(a, spans, evs)
    rb, eb = _run_steps(b, spans, evs)
    for i, ((la, ga), (lb, gb)) in enumerate(zip(ra, rb)):
        assert abs(la - lb) < 1e-6 * max(1.0, abs(lb)), (i, la, lb)
        assert float((ga - gb).abs().max()) <= 2e-5 * max(float(gb.abs().max()), 1e-3), i
    assert torch.allclose(a.tgn.memory.memory, b.tgn.memory.memory, rtol=1e-5, atol=1e-7)
    for x, y in zip(ea, eb):
        assert torch.equal(x[2], y[2]) and torch.allclose(x[3], y[3], rtol=1e-5, atol=1e-6)
    hb = a.make_host_batches(12000, 1, 256)[0]
    assert float(a.train_step_host(hb).item()) > 0

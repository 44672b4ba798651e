"""Model-level parity on the GPU: the drop-in TGN (overlay) against the reference's golden
vectors and against the CPU oracle, reading like the reference's own call sequence."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import load_golden, oracle_from_golden, rel_err, batch_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OVERLAY = os.path.join(ROOT, "pfotgnrec_b200", "overlay")
TOL = 1e-5          # fp32 contract of BASELINE.json north_star (max-norm relative error)
GTOL = 5e-5         # parameter gradients: sums over thousands of rows in a different order


@pytest.fixture(scope="module")
def overlay():
    sys.path.insert(0, OVERLAY)
    for m in [k for k in sys.modules if k.split(".")[0] in ("model", "modules", "utils")]:
        del sys.modules[m]
    import model.tgn as tgn_mod
    import utils.utils as utils_mod
    yield tgn_mod, utils_mod
    sys.path.remove(OVERLAY)


def build_tgn(tgn_mod, utils_mod, z, gemm_mode="fp32"):
    import types
    data = types.SimpleNamespace(sources=z["st_sources"], destinations=z["st_destinations"],
                                 edge_idxs=z["st_edge_idxs"], timestamps=z["st_timestamps"])
    nf = utils_mod.get_neighbor_finder(data, uniform=False, max_node_idx=int(z["st_n_nodes"]) - 1)
    torch.manual_seed(11)
    tgn = tgn_mod.TGN(neighbor_finder=nf, node_features=z["node_feat"], edge_features=z["st_edge_features"].copy(),
                      device=torch.device("cuda"), n_layers=int(z["cfg_n_layers"]), n_heads=2, dropout=0.0,
                      use_memory=bool(z["cfg_use_memory"]), message_dimension=100, memory_dimension=int(z["cfg_d"]),
                      memory_update_at_start=True, embedding_module_type=str(z["cfg_embedding"]),
                      message_function=str(z["cfg_msg_fn"]) if "cfg_msg_fn" in z else "identity",
                       aggregator_type=str(z["cfg_aggregator"]) if "cfg_aggregator" in z else "last",
                      memory_updater_type=str(z["cfg_updater"]), n_neighbors=int(z["cfg_n_neighbors"]),
                      mean_time_shift_src=z["cfg_shift"][0], std_time_shift_src=z["cfg_shift"][1],
                      mean_time_shift_dst=z["cfg_shift"][2], std_time_shift_dst=z["cfg_shift"][3],
                      use_destination_embedding_in_message=bool(z["cfg_dst_emb"]),
                      use_source_embedding_in_message=bool(z["cfg_src_emb"]) if "cfg_src_emb" in z else False,
                       dyrep=bool(z["cfg_dyrep"]), gemm_mode=gemm_mode)
    tgn = tgn.to(torch.device("cuda"))
    sd = tgn.state_dict()
    for k in z:
        if k.startswith("w_"):
            assert np.array_equal(sd[k[2:]].cpu().numpy(), z[k]), k     # same seed -> same initial weights
    return tgn


def check_against_vectors(tgn_mod, utils_mod, z, mode, tag, tol=TOL, gtol=GTOL):
    """The drop-in TGN on the batches recorded in `z` (vectors the reference produced on the same inputs, seed and
    initial weights): embeddings, BPR loss, every parameter gradient, memory, last_update and the pending-message
    table after each batch.  Returns the largest relative error seen per quantity."""
    tgn = build_tgn(tgn_mod, utils_mod, z, gemm_mode=mode).train()
    n = int(z["cfg_n_neighbors"])
    n_neg = int(z["cfg_n_neg"])
    names = [k for k, p in tgn.named_parameters() if p.requires_grad]
    worst = {"emb": 0.0, "loss": 0.0, "grad": 0.0, "memory": 0.0, "pend_msg": 0.0}
    for bi in range(int(z["cfg_n_batches"])):
        src, dst, extra, ts, ei = batch_inputs(z, bi)
        tgn.zero_grad(set_to_none=True)
        if len(extra) == 2:
            e_s, e_d, e_p, e_n = tgn.compute_temporal_embeddings_p(src, dst, extra[0], extra[1], ts, ei, n)
            outs = {"src": e_s, "dst": e_d, "ppos": e_p, "neg": e_n}
        else:
            e_s, e_d, e_n = tgn.compute_temporal_embeddings(src, dst, extra[0], ts, ei, n)
            e_p = e_d
            outs = {"src": e_s, "dst": e_d, "neg": e_n}
        for nm, e in outs.items():
            err = rel_err(e.detach().cpu().numpy(), z[f"b{bi}_emb_{nm}"])
            worst["emb"] = max(worst["emb"], err)
            assert err < tol, (tag, bi, nm, err)
        # BPR exactly as the reference script computes it (main.py:321-337), on the returned tensors
        bs = e_s.shape[0]
        s_ = e_s.view(bs, 1, -1)
        pos = torch.sum(s_ * e_p.view(bs, 1, -1), dim=2)
        neg = torch.matmul(s_, e_n.view(bs, n_neg, -1).transpose(1, 2)).squeeze()
        loss = -torch.mean(torch.log(torch.sigmoid(torch.mean(pos - neg, dim=1))))
        ref_loss = float(z[f"b{bi}_loss"])
        worst["loss"] = max(worst["loss"], abs(loss.item() - ref_loss) / max(1.0, abs(ref_loss)))
        assert abs(loss.item() - ref_loss) < tol * max(1.0, abs(ref_loss)), (tag, bi, loss.item(), ref_loss)
        if not loss.requires_grad:      # dyrep (main.py:386-387); identity embedding of an all-zero memory
            loss.requires_grad_()
        loss.backward()
        if tgn.use_memory:
            tgn.memory.detach_memory()
        params = dict(tgn.named_parameters())
        for k in names:
            ref = z[f"b{bi}_g_{k}"]
            g = params[k].grad
            g = np.zeros_like(ref) if g is None else g.cpu().numpy()
            scale = max(np.abs(ref).max(), 1e-3)
            worst["grad"] = max(worst["grad"], float(np.abs(g - ref).max() / scale))
            assert np.abs(g - ref).max() <= gtol * scale + 1e-7, (tag, bi, k, np.abs(g - ref).max(), scale)
        if tgn.use_memory:
            worst["memory"] = max(worst["memory"], rel_err(tgn.memory.memory.cpu().numpy(), z[f"b{bi}_memory"]))
            assert rel_err(tgn.memory.memory.cpu().numpy(), z[f"b{bi}_memory"]) < tol
            assert np.array_equal(tgn.memory.last_update.cpu().numpy(), z[f"b{bi}_last_update"])
            st = tgn.memory.state
            v = z[f"b{bi}_pend_valid"]
            assert np.array_equal(st.pend_valid.cpu().numpy().astype(bool), v)           # last-message selection: bit-exact
            assert np.array_equal(st.pend_ts.cpu().numpy()[v], z[f"b{bi}_pend_ts"][v])
            raw = z[f"b{bi}_pend_msg"].shape[1]
            err = rel_err(st.pend_msg.cpu().numpy()[v][:, :raw], z[f"b{bi}_pend_msg"][v])
            worst["pend_msg"] = max(worst["pend_msg"], err)
            assert err < tol
    return worst


GOLDEN_TAGS = ["ours", "ours_nbg", "tgn", "jodie", "dyrep", "tgat2", "tgat2x20", "mlp_mean", "srcemb", "gsum", "gsum2",
               "identity"]


@pytest.mark.parametrize("tag", GOLDEN_TAGS)
@pytest.mark.parametrize("mode", ["fp32", "simt"])
def test_drop_in_tgn_matches_reference_golden(overlay, tag, mode):
    """1e-5 contract in both exact GEMM modes: "fp32" = 3xTF32 on the tcgen05 tensor cores (default),
    "simt" = FFMA."""
    tgn_mod, utils_mod = overlay
    check_against_vectors(tgn_mod, utils_mod, load_golden(f"tgn_{tag}.npz"), mode, tag)


def test_larger_stream_vs_oracle(overlay):
    """PfoTGNRec config (d=64, n=10, heads 2) on a bigger synthetic stream, weights injected into the
    CPU oracle; several batches so that memory, last-wins messages and lazy updates interact."""
    tgn_mod, utils_mod = overlay
    import types
    from oracle.graph import AdjacencyOracle
    from oracle.tgn import TGNOracle, bpr_loss
    from pfotgnrec_b200.synth import make_stream
    st = make_stream(n_users=400, n_items=50, n_events=3000, n_days=30, seed=4, ts_mode="small", with_prices=False)
    rng = np.random.default_rng(0)
    d, B, n = 64, 200, 10
    node_feat = rng.random((st.n_nodes, d))
    data = types.SimpleNamespace(sources=st.sources, destinations=st.destinations, edge_idxs=st.edge_idxs,
                                 timestamps=st.timestamps)
    nf = utils_mod.get_neighbor_finder(data, uniform=False, max_node_idx=st.n_nodes - 1)
    torch.manual_seed(3)
    tgn = tgn_mod.TGN(neighbor_finder=nf, node_features=node_feat, edge_features=st.edge_features.copy(),
                      device=torch.device("cuda"), n_layers=1, n_heads=2, dropout=0.0, use_memory=True,
                      message_dimension=100, memory_dimension=d, embedding_module_type="graph_attention",
                      message_function="identity", aggregator_type="last", memory_updater_type="gru",
                      n_neighbors=n).to("cuda").train()
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in tgn.named_parameters()}
    adj = AdjacencyOracle(st.sources, st.destinations, st.edge_idxs, st.timestamps, n_nodes=st.n_nodes)
    orc = TGNOracle(p, adj, node_feat, st.edge_features, n_layers=1, n_heads=2)
    for bi in range(8):
        sl = slice(bi * B, (bi + 1) * B)
        ppos = rng.integers(st.n_users + 1, st.n_nodes, size=B)
        pneg = rng.integers(st.n_users + 1, st.n_nodes, size=3 * B)
        args = (st.sources[sl], st.destinations[sl])
        tgn.zero_grad(set_to_none=True)
        for v in p.values():
            v.grad = None
        e = tgn.compute_temporal_embeddings_p(*args, ppos, pneg, st.timestamps[sl], st.edge_idxs[sl], n)
        o = orc.compute_temporal_embeddings(*args, [ppos, pneg], st.timestamps[sl], st.edge_idxs[sl], n)
        for a, b in zip(e, o):
            assert rel_err(a.detach().cpu().numpy(), b.detach().numpy()) < TOL, bi
        la = bpr_loss(e[0], e[2], e[3])
        lb = bpr_loss(o[0], o[2], o[3])
        assert abs(la.item() - lb.item()) < TOL
        la.backward()
        lb.backward()
        for k, v in tgn.named_parameters():
            if p[k].grad is None:
                continue
            ref = p[k].grad.numpy()
            got = v.grad.cpu().numpy() if v.grad is not None else np.zeros_like(ref)
            scale = max(np.abs(ref).max(), 1e-3)
            assert np.abs(got - ref).max() <= GTOL * scale + 1e-7, (bi, k)
        assert rel_err(tgn.memory.memory.cpu().numpy(), orc.memory.numpy()) < TOL
        assert np.array_equal(tgn.memory.state.pend_valid.cpu().numpy().astype(bool), orc.pend_valid.numpy())


@pytest.mark.parametrize("tag", ["ours", "jodie"])
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_fast_gemm_modes_within_2e_2(overlay, tag, mode):
    """gemm_mode='tf32' (single-pass TF32, TMA-fed tcgen05) and 'bf16' (tcgen05): the 2e-2 contract of the
    north_star on embeddings, loss, memory."""
    tgn_mod, utils_mod = overlay
    z = load_golden(f"tgn_{tag}.npz")
    tgn = build_tgn(tgn_mod, utils_mod, z, gemm_mode=mode).train()
    n, n_neg = int(z["cfg_n_neighbors"]), int(z["cfg_n_neg"])
    from pfotgnrec_b200.trainer import bpr_loss
    for bi in range(int(z["cfg_n_batches"])):
        src, dst, extra, ts, ei = batch_inputs(z, bi)
        tgn.zero_grad(set_to_none=True)
        if len(extra) == 2:
            e_s, e_d, e_p, e_n = tgn.compute_temporal_embeddings_p(src, dst, extra[0], extra[1], ts, ei, n)
        else:
            e_s, e_d, e_n = tgn.compute_temporal_embeddings(src, dst, extra[0], ts, ei, n)
            e_p = e_d
        for nm, e in (("src", e_s), ("dst", e_d), ("neg", e_n)):
            assert rel_err(e.detach().cpu().numpy(), z[f"b{bi}_emb_{nm}"]) < 2e-2, (tag, bi, nm)
        loss = bpr_loss(e_s, e_p, e_n)
        ref_loss = float(z[f"b{bi}_loss"])
        assert abs(loss.item() - ref_loss) < 2e-2 * max(1.0, abs(ref_loss))
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in tgn.parameters() if p.grad is not None)
        assert rel_err(tgn.memory.memory.cpu().numpy(), z[f"b{bi}_memory"]) < 2e-2
        assert np.array_equal(tgn.memory.state.pend_valid.cpu().numpy().astype(bool), z[f"b{bi}_pend_valid"])


def test_reference_checkpoint_continues_on_the_drop_in(overlay):
    """tests/golden/state_dict.npz: a state_dict written by the unmodified reference after two Adam steps, loaded into
    a fresh reference model that then ran batch 3 in eval mode.  The drop-in loads the same dictionary (strict) and has
    to reproduce that batch: embeddings, memory and last_update afterwards."""
    import types
    from test_overlay_host import build_for_state_dict, _state_dict_case
    tgn_mod, utils_mod = overlay
    z, sd = _state_dict_case()
    data = types.SimpleNamespace(sources=z["st_sources"], destinations=z["st_destinations"],
                                 edge_idxs=z["st_edge_idxs"], timestamps=z["st_timestamps"])
    nf = utils_mod.get_neighbor_finder(data, uniform=False, max_node_idx=int(z["st_n_nodes"]) - 1)
    tgn = build_for_state_dict(tgn_mod, z, "cuda", nf=nf).to(torch.device("cuda"))
    tgn.load_state_dict(sd, strict=True)
    tgn.eval()
    B, n = int(z["cfg_B"]), int(z["cfg_n"])
    sl = slice(2 * B, 3 * B)
    with torch.no_grad():
        e_s, e_d, e_n = tgn.compute_temporal_embeddings(z["st_sources"][sl], z["st_destinations"][sl], z["neg"],
                                                        z["st_timestamps"][sl], z["st_edge_idxs"][sl], n)
    for nm, e in (("src", e_s), ("dst", e_d), ("neg", e_n)):
        assert rel_err(e.cpu().numpy(), z[f"emb_{nm}"]) < TOL, nm
    assert rel_err(tgn.memory.memory.cpu().numpy(), z["memory_after"]) < TOL
    assert np.array_equal(tgn.memory.last_update.cpu().numpy(), z["last_update_after"])


def test_backup_restore_memory_round_trip(overlay):
    """modules/memory.py:48-60: `backup_memory()` returns (memory, last_update, messages) clones; after more batches,
    `restore_memory(backup)` brings memory, last_update AND the pending messages back, so the next batch reproduces
    what it produced the first time (main.py:418 takes the backup after validation)."""
    tgn_mod, utils_mod = overlay
    z = load_golden("tgn_ours.npz")
    tgn = build_tgn(tgn_mod, utils_mod, z).eval()
    n = int(z["cfg_n_neighbors"])

    def run(bi):
        src, dst, extra, ts, ei = batch_inputs(z, bi)
        with torch.no_grad():
            return [e.clone() for e in tgn.compute_temporal_embeddings_p(src, dst, extra[0], extra[1], ts, ei, n)]

    run(0); run(1)
    backup = tgn.memory.backup_memory()
    assert len(backup) == 3                                         # (memory, last_update, messages)
    mem_at_backup = tgn.memory.memory.clone()
    first = run(2)
    run(3)
    assert not torch.equal(tgn.memory.memory, mem_at_backup)
    backup[0].add_(0.0)                                             # the backup is a clone, not a view of the state
    tgn.memory.restore_memory(backup)
    assert torch.equal(tgn.memory.memory, mem_at_backup)
    assert torch.equal(tgn.memory.last_update, backup[1])
    again = run(2)
    for a, b in zip(first, again):
        assert torch.equal(a, b)
    # the state after restore + batch 2 equals the reference's golden after batch 2 (eval mode == dropout 0 here)
    assert rel_err(tgn.memory.memory.cpu().numpy(), z["b2_memory"]) < TOL
    assert np.array_equal(tgn.memory.state.pend_valid.cpu().numpy().astype(bool), z["b2_pend_valid"])


def test_standalone_sub_module_forwards():
    """Reference-derived scripts may call a sub-module directly: MergeLayer (utils/utils.py:14-17), the MLP message
    function (modules/message_function.py:23-26) and TimeEncode (model/time_encoding.py:17-25) evaluate on their own
    through the library, against the same arithmetic in fp64 torch."""
    from pfotgnrec_b200.containers import MergeLayer, MLPMessageFunction
    torch.manual_seed(0)
    ml = MergeLayer(128, 64, 64, 64).cuda()
    x1, x2 = torch.randn(300, 128, device="cuda"), torch.randn(300, 64, device="cuda")
    ref = torch.relu(torch.cat([x1, x2], 1).double() @ ml.fc1.weight.double().T + ml.fc1.bias.double())
    ref = ref @ ml.fc2.weight.double().T + ml.fc2.bias.double()
    assert rel_err(ml(x1, x2).cpu().numpy(), ref.detach().cpu().numpy()) < TOL
    mf = MLPMessageFunction(193, 100).cuda()
    raw = torch.randn(257, 193, device="cuda")
    ref = torch.relu(raw.double() @ mf.mlp[0].weight.double().T + mf.mlp[0].bias.double())
    ref = ref @ mf.mlp[2].weight.double().T + mf.mlp[2].bias.double()
    assert rel_err(mf.compute_message(raw).cpu().numpy(), ref.detach().cpu().numpy()) < TOL


@pytest.mark.parametrize("tag", ["ours", "jodie", "mlp_mean"])
def test_merged_cell_gemm_matches_reference_golden(overlay, tag):
    """ModelConfig.cell_gemm = "merged" (one contraction over [message | memory] with the block weight of pfo_pack_cell,
    gradients unpacked by pfo_unpack_cell_grads; opt-in, see engine.ModelConfig) under the same 1e-5 / 5e-5 contract:
    GRU, RNN and the MLP message function in front of the cell."""
    tgn_mod, utils_mod = overlay
    z = load_golden(f"tgn_{tag}.npz")
    orig = tgn_mod.TGN._get_engine

    def merged_engine(self):
        self._cfg.cell_gemm = "merged"
        return orig(self)

    tgn_mod.TGN._get_engine = merged_engine
    try:
        check_against_vectors(tgn_mod, utils_mod, z, "fp32", tag + "/merged")
    finally:
        tgn_mod.TGN._get_engine = orig

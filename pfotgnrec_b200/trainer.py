"""Training / evaluation driver over a synthetic stream: the loop of reference main.py:160-394
(train) and evaluation.py:63-138 (scoring + ranking), with the MV-selection block and the BPR
loss -- inline script code in the reference -- behind entry points (`MVSelector.select`,
`bpr_loss`).  Everything per batch runs on the device; the host only launches.
"""
from __future__ import annotations

import importlib
import os
import sys
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import ptr
from .engine import ModelConfig
from .graph import TemporalCSR, NeighborFinder
from .sampler import CandidateSampler, MVSelector
from .evalmetrics import EvalMetricBlock
from .synth import Stream, log_returns

_OVERLAY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay")


def load_overlay():
    """Import the drop-in packages (model.tgn, utils.utils, ...) from pfotgnrec_b200/overlay."""
    if _OVERLAY not in sys.path:
        sys.path.insert(0, _OVERLAY)
    return importlib.import_module("model.tgn"), importlib.import_module("utils.utils")


class FlatAdam:
    """torch.optim.Adam(lr) (main.py:123; betas 0.9 / 0.999, eps 1e-8, no weight decay) over one flat parameter buffer:
    a single `pfo_adam_flat` launch per step, step count on the device (graph replays keep counting)."""

    def __init__(self, flat_params, flat_grads, lr, betas=(0.9, 0.999), eps=1e-8):
        self.p, self.g = flat_params, flat_grads
        self.m, self.v = torch.zeros_like(flat_params), torch.zeros_like(flat_params)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.t = torch.zeros(1, dtype=torch.int32, device=flat_params.device)

    def step(self):
        self.t.add_(1)
        _lib.call("pfo_adam_flat", ptr(self.p), ptr(self.g), ptr(self.m), ptr(self.v), self.p.numel(), self.lr,
                  self.betas[0], self.betas[1], self.eps, ptr(self.t))

    def zero_grad(self, set_to_none=False):
        self.g.zero_()

    def state_dict(self):
        return {"exp_avg": self.m, "exp_avg_sq": self.v, "step": self.t, "lr": self.lr, "betas": self.betas, "eps": self.eps}


class _BPR(torch.autograd.Function):
    """-mean log sigmoid(mean_k(<u,p> - <u,n_k>)) with its gradient from one fused kernel (K6).  `grad_scale`
    multiplies the gradients inside the kernel (the data-parallel trainer's 1 / world); with `unit_upstream` the
    caller promises to call `.backward()` on the returned loss itself, so the three gradient tensors are handed to
    autograd as they are instead of being multiplied by the upstream 1.0 (three elementwise launches per step)."""

    @staticmethod
    def forward(ctx, eu, ep, en, ws, grad_scale=1.0, unit_upstream=False):
        B, d = eu.shape
        k = en.shape[0] // B
        eu, ep, en = eu.contiguous(), ep.contiguous(), en.contiguous()
        loss = torch.empty(1, device=eu.device)
        need = any(ctx.needs_input_grad[:3])
        du = torch.empty_like(eu) if need else None
        dp = torch.empty_like(ep) if need else None
        dn = torch.empty_like(en) if need else None
        _lib.call("pfo_bpr", ptr(eu), ptr(ep), ptr(en), B, k, d, ptr(du), ptr(dp), ptr(dn), ptr(loss), float(grad_scale),
                  ptr(ws))
        ctx.grads, ctx.unit_upstream = (du, dp, dn), bool(unit_upstream)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        du, dp, dn = ctx.grads
        if ctx.unit_upstream:
            return du, dp, dn, None, None, None
        return du * g, dp * g, dn * g, None, None, None


class _BPRPacked(torch.autograd.Function):
    """The training step's BPR over the PACKED embedding tensor [Q, d] (row blocks [src | dst | (p_pos) | negatives]):
    the fused kernel reads its three operands as row blocks and writes their gradients into the matching blocks of
    one [Q, d] buffer, which is the gradient of the packed tensor -- no split / zero-fill / concatenation in autograd.
    The caller calls `.backward()` on the returned loss itself (upstream gradient 1; `grad_scale` applied in the kernel)."""

    @staticmethod
    def forward(ctx, emb, B, pos_block, neg_block, k, ws, grad_scale):
        d = emb.shape[1]
        emb = emb.contiguous()
        row = lambda t, blk: t.data_ptr() + blk * B * d * t.element_size()
        loss = torch.empty(1, device=emb.device)
        dEmb = None
        if ctx.needs_input_grad[0]:
            dEmb = torch.empty_like(emb)
            for blk in range(neg_block):                 # blocks the loss does not read (dst in the PfoTGNRec step)
                if blk not in (0, pos_block):
                    dEmb[blk * B:(blk + 1) * B].zero_()
        dptr = (lambda blk: row(dEmb, blk)) if dEmb is not None else (lambda blk: None)
        _lib.call("pfo_bpr", row(emb, 0), row(emb, pos_block), row(emb, neg_block), B, k, d,
                  dptr(0), dptr(pos_block), dptr(neg_block), ptr(loss), float(grad_scale), ptr(ws))
        ctx.dEmb = dEmb
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        return ctx.dEmb, None, None, None, None, None, None


def bpr_loss(e_u, e_pos, e_neg, workspace=None):
    """BPR loss of reference main.py:321-337 (e_neg is [B*k, d], interaction-major)."""
    if workspace is None:
        workspace = torch.empty(1024, device=e_u.device)
    return _BPR.apply(e_u, e_pos, e_neg, workspace)


@dataclass
class TrainConfig:
    model: str = "ours"          # ours | tgn | jodie | dyrep | tgat   (reference main.py:63-74)
    bs: int = 512
    n_neighbors: int = 10
    n_layers: int = 1
    d: int = 64
    n_heads: int = 2
    dropout: float = 0.0
    lr: float = 1e-4
    num_negatives: int = 20      # candidates per interaction for MV selection
    p_pos_num: int = 1
    p_neg_num: int = 3
    gamma: float = 2.0
    lambda_mv: float = 0.5
    seed: int = 0
    gemm_mode: str = "fp32"
    cuda_graph: bool = True      # replay the whole step (sampling .. Adam) as one CUDA graph per batch size


class StreamOnDevice:
    """The interaction stream resident in HBM (columns of reference utils/data.py `Data`)."""

    def __init__(self, st: Stream, device):
        dev = torch.device(device)
        self.n_events = st.n_events
        self.src = torch.as_tensor(st.sources.astype(np.int32), device=dev)
        self.dst = torch.as_tensor(st.destinations.astype(np.int32), device=dev)
        self.ts = torch.as_tensor(st.timestamps, device=dev)
        self.eidx = torch.as_tensor(st.edge_idxs.astype(np.int32), device=dev)
        self.ev = torch.as_tensor(st.edge_idxs.astype(np.int64), device=dev)
        self.day = torch.as_tensor(st.day_idx.astype(np.int32), device=dev)
        self.port_ptr = torch.as_tensor(st.port_ptr.astype(np.int64), device=dev)
        pi = st.port_items if st.port_items.size else np.zeros(1, np.int32)
        self.port_items = torch.as_tensor(pi.astype(np.int32), device=dev)
        self.port_items_as_item_ids = self.port_items + (st.n_users + 1)


def time_statistics(sources, destinations, timestamps):
    """reference utils/data.py:75-99 (compute_time_statistics) without the Python loop over events: per-node
    inter-event times through one stable sort by node id.  The mean / std run over the diffs in stream order,
    like the reference's lists, so the fp64 sums round identically."""
    out = []
    n = len(sources)
    for ids in (np.asarray(sources), np.asarray(destinations)):
        order = np.lexsort((np.arange(n), ids))
        t = np.asarray(timestamps, dtype=np.float64)[order]
        first = np.r_[True, ids[order][1:] != ids[order][:-1]]
        diff = np.empty(n, dtype=np.float64)
        diff[order] = np.where(first, t, t - np.r_[0.0, t[:-1]])
        out += [float(np.mean(diff)), float(np.std(diff))]
    return out


class _StepGraph:
    """Static input buffers, the captured graph and its output for one batch size."""

    def __init__(self, static):
        self.static, self.graph, self.loss, self.eager_steps, self.launches = static, None, None, 0, 0
        self.evals = {}              # (n_items, with state batch) -> _StepGraph of the evaluation step


class PfoTrainer:
    """Builds the model on a synthetic stream and steps it (train or evaluate)."""

    def __init__(self, st: Stream, tc: TrainConfig, device="cuda", train_frac_mask=None):
        self.st, self.tc, self.device = st, tc, _lib.use_device(device)
        if tc.p_pos_num != 1:
            # the fused BPR kernel (and the packed row blocks [src | dst | p_pos | p_neg]) address ONE p_pos row per
            # interaction; main.py:327-331 would average over p_pos_num rows -- refuse instead of mis-addressing
            raise NotImplementedError("p_pos_num != 1: the fused BPR path takes one MV-selected positive per interaction")
        self.epoch = 0               # folded into the key of the candidate / negative streams (fresh draws per epoch)
        tgn_mod, _ = load_overlay()
        info = self._prepare_stream(train_frac_mask)         # split, device columns, feature tables, item universes
        self._build_finders(info["train_index"])
        kw = dict(memory_updater_type="gru", embedding_module_type="graph_attention", use_memory=True,
                  dyrep=False, use_destination_embedding_in_message=False)
        if tc.model == "jodie":
            kw.update(memory_updater_type="rnn", embedding_module_type="time")
        elif tc.model == "dyrep":
            kw.update(memory_updater_type="rnn", dyrep=True, use_destination_embedding_in_message=True)
        elif tc.model == "tgat":
            kw.update(use_memory=False)
        ms, ss, md, sd = info["time_statistics"]
        torch.manual_seed(tc.seed)
        self.tgn = tgn_mod.TGN(neighbor_finder=self.nf_train, node_features=info["node_feat"],
                               edge_features=info["edge_feat"], device=self.device, n_layers=tc.n_layers,
                               n_heads=tc.n_heads, dropout=tc.dropout, message_dimension=100,
                               memory_dimension=tc.d, memory_update_at_start=True, message_function="identity",
                               aggregator_type="last", n_neighbors=tc.n_neighbors,
                               mean_time_shift_src=ms, std_time_shift_src=ss, mean_time_shift_dst=md,
                               std_time_shift_dst=sd, use_source_embedding_in_message=False,
                               gemm_mode=tc.gemm_mode, **self._tgn_extra(), **kw).to(self.device)
        self._bind_engine()
        if os.environ.get("PFO_STORE_OVERLAP", "1") != "0" and type(self) is PfoTrainer:
            # every training step of this loop runs its backward right after the forward: the state update of the
            # forward pass may leave the main stream (engine.TGNEngine.store_state)
            self.tgn._get_engine().overlap_store = True
        self._setup_optimizer()
        self._graphs = {}            # batch size -> _StepGraph
        universe_items = info["universe_train"]
        self.universe_items = universe_items
        self.mv = None
        if tc.model == "ours":
            self.mv = MVSelector(log_returns(st.prices_future), universe_items - st.n_users - 1, st.n_users,
                                 gamma=tc.gamma, lam=tc.lambda_mv, n_candidates=tc.num_negatives,
                                 n_pos=tc.p_pos_num, n_neg=tc.p_neg_num, seed=tc.seed, device=device)
        self.neg_sampler = CandidateSampler(universe_items, device=device)
        self.eval_sampler = CandidateSampler(info["universe_all"], device=device)
        self.bpr_ws = torch.empty(1024, device=self.device)
        # evaluation metric block (reference evaluation.py:127-258): in-sample (past) / out-of-sample (future)
        # daily log-returns and the running sums of the per-interaction metrics live on the device
        self.metrics = None
        if st.prices_past.shape[0] == len(st.day_keys) and st.prices_past.shape[1] == st.n_items:
            self.metrics = EvalMetricBlock(log_returns(st.prices_past), log_returns(st.prices_future),
                                           st.n_users + 1, device=self.device)

    def _setup_optimizer(self):
        """Adam (main.py:123).  When every operand of the step is a reference parameter as it is, the parameters are
        re-homed into ONE flat buffer in the engine's operand order, the step's backward writes their gradients
        straight into the flat gradient buffer `gflat` (engine.grad_sink; `.grad` of each parameter is a view of it, and
        it is the bucket the multi-GPU trainers all-reduce), and the optimiser step is one launch (`pfo_adam_flat`).
        Otherwise (graph_sum derives its operands with torch ops) torch's fused Adam runs on the separate tensors."""
        tc, eng = self.tc, self.tgn._get_engine()
        self.gflat, self.flat_params = None, None
        if not eng.flat_layout_ok():
            self.opt = torch.optim.Adam(self.tgn.parameters(), lr=tc.lr, fused=True, capturable=True)
            return
        named = dict(self.tgn.named_parameters(remove_duplicate=False))
        plist = [named[k] for k in eng.param_names()]
        sizes = [p.numel() for p in plist]
        if any(n % 4 for n in sizes):                     # slices must stay 16-byte aligned for the GEMM operand loads
            self.opt = torch.optim.Adam(self.tgn.parameters(), lr=tc.lr, fused=True, capturable=True)
            return
        flat = torch.cat([p.detach().reshape(-1) for p in plist]).contiguous()
        self.gflat = torch.zeros_like(flat)
        for p, w, g in zip(plist, flat.split(sizes), self.gflat.split(sizes)):
            p.data = w.view_as(p)                         # same values, now a slice of the flat buffer
            p.grad = g.view_as(p)
        self.flat_params = flat
        eng.grad_sink = self.gflat
        self.opt = FlatAdam(flat, self.gflat, lr=tc.lr)

    def _prepare_stream(self, train_frac_mask=None):
        """Everything the constructor derives from the interaction stream (host `synth.Stream`): the chronological
        split (utils/data.py:27,48-50), the columns resident on the device, node features ~ U(0,1) (main.py:9,87), the
        edge features, the time statistics (utils/data.py:75-99) and the item universes of the two samplers."""
        st, tc = self.st, self.tc
        train_mask, val_mask, test_mask = st.split()
        if train_frac_mask is not None:
            train_mask = train_frac_mask
        self.masks = (train_mask, val_mask, test_mask)
        self.n_train = int(train_mask.sum())
        tr = np.nonzero(train_mask)[0]
        self.dev_stream = StreamOnDevice(st, self.device)
        return dict(train_index=tr, node_feat=np.random.RandomState(0).rand(st.n_nodes, tc.d),
                    edge_feat=st.edge_features.copy(), time_statistics=self._time_statistics(),
                    universe_train=np.unique(st.destinations[tr]), universe_all=np.unique(st.destinations))

    @property
    def eval_acc(self):
        return self.metrics.acc

    # ---- hooks of the node-sharded subclass (pfotgnrec_b200/dist.py)
    def _build_finders(self, tr):
        """Train-split and full adjacency (main.py:95-96) as device CSRs + the K1 finders on top."""
        st, tc = self.st, self.tc
        self.csr_train = TemporalCSR(st.sources[tr], st.destinations[tr], st.edge_idxs[tr], st.timestamps[tr],
                                     n_nodes=st.n_nodes, device=self.device)
        self.csr_full = TemporalCSR(st.sources, st.destinations, st.edge_idxs, st.timestamps,
                                    n_nodes=st.n_nodes, device=self.device)
        uniform = tc.model == "tgat"
        self.nf_train = NeighborFinder(self.csr_train, uniform=uniform, seed=tc.seed)
        self.nf_full = NeighborFinder(self.csr_full, uniform=uniform, seed=tc.seed)

    def _tgn_extra(self):
        return {}

    def _bind_engine(self):
        pass

    def _time_statistics(self):
        return time_statistics(self.st.sources, self.st.destinations, self.st.timestamps)

    # ------------------------------------------------------------------ one training step
    def _ev_offset(self):
        """Offset of the Philox stream ids in the current epoch.  The reference draws fresh candidates every epoch
        (np.random.choice from the unseeded global stream, main.py:194-195, utils/utils.py:103-111); here a draw is
        keyed by (seed, stream id of the interaction, slot), so the epoch is folded into the id: id = edge idx +
        epoch * (n_events + 1).  It rides in the batch's `ev` column -- a graph replay reads it from the static
        buffer, nothing is baked into the captured kernels."""
        return int(self.epoch) * (int(self.st.n_events) + 1)

    def _batch(self, s, e):
        D = self.dev_stream
        off = self._ev_offset()
        return dict(src=D.src[s:e], dst=D.dst[s:e], ts=D.ts[s:e], eidx=D.eidx[s:e],
                    ev=D.ev[s:e] + off if off else D.ev[s:e], day=D.day[s:e], port_ptr=D.port_ptr[s:e + 1])

    # ---- one staging buffer per batch.  The eight columns of a batch live back to back in ONE byte buffer (16-byte
    # aligned fields, the variable-length portfolio entries last), on the host (pinned) and in the static device
    # buffers of the captured step alike, so a host batch reaches the device as ONE cudaMemcpyAsync instead of eight
    # (each costs ~3 us of launch on the host, exposed when the caller reads the loss back every step).
    _PACK_FIELDS = (("src", torch.int32, 0), ("dst", torch.int32, 0), ("eidx", torch.int32, 0), ("day", torch.int32, 0),
                    ("ts", torch.float64, 0), ("ev", torch.int64, 0), ("port_ptr", torch.int64, 1))

    @classmethod
    def _pack_layout(cls, B, n_port):
        """name -> (byte offset, elements, dtype) of a B-interaction batch with n_port portfolio entries; total bytes."""
        lay, off = {}, 0
        for name, dt, extra in cls._PACK_FIELDS + (("port_items", torch.int32, None),):
            n = n_port if extra is None else B + extra
            lay[name] = (off, n, dt)
            off = (off + n * torch.empty((), dtype=dt).element_size() + 15) // 16 * 16
        return lay, off

    @classmethod
    def _packed_views(cls, buf, B, n_port):
        lay, _ = cls._pack_layout(B, n_port)
        return {k: buf[o:o + n * torch.empty((), dtype=dt).element_size()].view(dt) for k, (o, n, dt) in lay.items()}

    def make_host_batches(self, start, count, bs):
        """Pinned host copies of `count` consecutive batches, as a caller holding numpy data passes them: the columns
        are views of one pinned staging buffer per batch (`_packed`)."""
        st, out = self.st, []
        for i in range(count):
            s, e = start + i * bs, start + (i + 1) * bs
            pp = st.port_ptr[s:e + 1]
            cols = dict(src=st.sources[s:e].astype(np.int32), dst=st.destinations[s:e].astype(np.int32),
                        ts=st.timestamps[s:e].astype(np.float64), eidx=st.edge_idxs[s:e].astype(np.int32),
                        ev=st.edge_idxs[s:e].astype(np.int64) + self._ev_offset(), day=st.day_idx[s:e].astype(np.int32),
                        port_ptr=(pp - pp[0]).astype(np.int64),
                        port_items=np.r_[st.port_items[pp[0]:pp[-1]], 0].astype(np.int32))
            out.append(self._pack_host_batch(cols))
        return out

    def _pack_host_batch(self, cols):
        """dict of numpy / CPU-tensor columns -> the same columns as views of one pinned byte buffer."""
        B, n_port = int(cols["src"].shape[0]), int(cols["port_items"].shape[0])
        _, nbytes = self._pack_layout(B, n_port)
        buf = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        hb = self._packed_views(buf, B, n_port)
        for k, v in hb.items():
            v.copy_(torch.as_tensor(np.ascontiguousarray(cols[k])) if isinstance(cols[k], np.ndarray) else cols[k])
        hb["_packed"] = buf
        hb["nbytes"] = nbytes
        return hb

    def _copy_host_batch(self, sg, hb):
        """Host batch -> the static device buffers of a captured step: one copy when both sides are packed alike."""
        x = sg.static
        buf = hb.get("_packed")
        if buf is not None and "_packed" in x and hb["src"].shape[0] == x["src"].shape[0] \
                and buf.numel() <= x["_packed"].numel():
            x["_packed"][:buf.numel()].copy_(buf, non_blocking=True)
            return
        for k, v in hb.items():
            if k in ("nbytes", "_packed", "_global"):
                continue
            dst = x[k]
            (dst[:v.shape[0]] if k == "port_items" else dst).copy_(v, non_blocking=True)

    def train_step_host(self, hb):
        """One step from HOST buffers: host->device copy of the batch, then `train_step`."""
        B = hb["src"].shape[0]
        if self._graph_ok(B):
            sg = self._step_graph(B)
            self._copy_host_batch(sg, hb)
            return self._run_graphed(sg)
        b = {k: v.to(self.device, non_blocking=True) for k, v in hb.items() if k not in ("nbytes", "_packed", "_global")}
        return self.train_step(0, 0, batch=b)

    # ---- CUDA-graph replay.  Every shape of the step is a function of the batch size alone (the unique-node
    # count lives on the device, kernels read it there), so one graph per batch size covers the stream.
    def _graph_ok(self, B):
        return bool(self.tc.cuda_graph) and self.device.type == "cuda"

    def _step_graph(self, B):
        sg = self._graphs.get(B)
        if sg is None:
            cap = self._port_capacity(B)
            _, nbytes = self._pack_layout(B, cap)
            buf = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            static = self._packed_views(buf, B, cap)         # the columns are views of one buffer (see _pack_layout)
            static["_packed"] = buf
            sg = self._graphs[B] = _StepGraph(static)
        return sg

    def _port_capacity(self, B):
        """Entries of the static portfolio buffer of a B-interaction batch (largest portfolio x B, + 1 spare)."""
        pp = self.st.port_ptr
        return int(np.max(np.diff(pp))) * B + 1 if pp.size > 1 else 1

    def _fill_static(self, sg, s, e):
        D, st = self.dev_stream, self.st
        x = sg.static
        x["src"].copy_(D.src[s:e]); x["dst"].copy_(D.dst[s:e]); x["ts"].copy_(D.ts[s:e])
        x["eidx"].copy_(D.eidx[s:e]); x["day"].copy_(D.day[s:e])
        torch.add(D.ev[s:e], self._ev_offset(), out=x["ev"])
        p0, p1 = int(st.port_ptr[s]), int(st.port_ptr[e])          # host copy of the CSR: no device sync
        torch.sub(D.port_ptr[s:e + 1], p0, out=x["port_ptr"])
        if p1 > p0:
            x["port_items"][:p1 - p0].copy_(D.port_items[p0:p1])

    def _run_graphed(self, sg):
        """The first steps at a batch size run eagerly on the static buffers (they are real training steps and
        warm every lazy allocation); the next one is captured (capture launches nothing) and replayed."""
        if sg.graph is None:
            if sg.eager_steps < 2:
                sg.eager_steps += 1
                return self._step_body(sg.static)
            torch.cuda.synchronize(self.device)
            self._zero_grads()
            g = torch.cuda.CUDAGraph()
            launches0 = _lib.LAUNCHES
            with torch.cuda.graph(g, **self._capture_kw):
                sg.loss = self._fwd_bwd(sg.static) if self._graph_tail_eager else self._step_body(sg.static)
            sg.launches = _lib.LAUNCHES - launches0
            sg.graph = g
        sg.graph.replay()
        _lib.LAUNCHES += sg.launches
        if self._graph_tail_eager:                   # gradient exchange + optimiser outside the graph
            self._finish()
        return sg.loss

    def train_step(self, s, e, batch=None):
        """Events [s, e) of the stream: MV selection (or uniform negatives) -> embeddings -> BPR ->
        backward -> Adam (reference main.py:179-394).  Returns the loss tensor (no host sync; in CUDA-graph
        mode it is the graph's output buffer, overwritten by the next step)."""
        if batch is None and self._graph_ok(e - s) and e - s > 0 and e <= self.st.n_events:
            sg = self._step_graph(e - s)
            self._fill_static(sg, s, e)
            return self._run_graphed(sg)
        return self._step_body(batch if batch is not None else self._batch(s, e))

    def _step_body(self, b):
        loss = self._fwd_bwd(b)
        self._finish()
        return loss

    def _finish(self):
        with _lib.nvtx_range("gradient exchange + Adam"):
            self._reduce_grads()
            self.opt.step()

    def _fwd_bwd(self, b):
        tc, D = self.tc, self.dev_stream
        tgn = self.tgn.train()
        eng = tgn._get_engine()
        eng.nf = self.nf_train
        self._zero_grads()
        params = tgn._params()
        B = b["src"].shape[0]
        port_items = b["port_items"] if "port_items" in b else D.port_items
        sb = b.get("state")                          # replicated data-parallel mode: the global batch advances the state
        if tc.model == "ours":
            with _lib.nvtx_range("K5 candidate sampling + MV selection"):
                p_pos, p_neg = self.mv.select(b["ev"], b["day"], b["dst"], b["port_ptr"], port_items)
            emb = eng.compute_temporal_embeddings(params, b["src"], b["dst"], [p_pos, p_neg], b["ts"], b["eidx"],
                                                  tc.n_neighbors, train=True, state_batch=sb, packed=True)
            pos_block, neg_block = 2, 3                  # rows [src | dst | p_pos | p_neg]
        else:
            held = port_items + (self.st.n_users + 1) if "port_items" in b else D.port_items_as_item_ids
            neg = self.neg_sampler.sample(b["ev"], b["port_ptr"], held, tc.p_neg_num, seed=tc.seed).reshape(-1)
            emb = eng.compute_temporal_embeddings(params, b["src"], b["dst"], [neg], b["ts"], b["eidx"],
                                                  tc.n_neighbors, train=True, state_batch=sb, packed=True)
            pos_block, neg_block = 1, 2                  # rows [src | dst | negatives]
        with _lib.nvtx_range("K6 BPR forward + backward"):
            loss = _BPRPacked.apply(emb, B, pos_block, neg_block, tc.p_neg_num, self.bpr_ws, self._loss_scale)
        with _lib.nvtx_range("backward"):
            loss.backward()
        return loss.detach()

    # hooks of the data-parallel subclass
    _loss_scale = 1.0
    _graph_tail_eager = False
    _capture_kw = {}             # multi-rank trainers capture NCCL collectives: capture_error_mode="thread_local"

    def _zero_grads(self):
        if self.gflat is None:
            self.opt.zero_grad(set_to_none=True)          # flat layout: the backward zero-fills its gradient buffer itself

    def _reduce_grads(self):
        pass

    # ------------------------------------------------------------------ one evaluation step
    @torch.no_grad()
    def eval_step(self, s, e, n_items=None, batch=None, state_batch=None, ev_base=None):
        """Interactions [s, e): N_ITEMS candidates per interaction (seed 2024), embeddings, scores, rank of
        the true item and top-5 (reference evaluation.py:84-115,134-138), then the metric block (:127-207) into
        `per_event` float64[B,18] and the running sums `self.eval_acc`.  Advances memory like the
        reference does.  Returns (pos_rank int32[B], top5 int32[B,5], candidates int32[B,N], scores, per_event); in
        CUDA-graph mode these are the graph's output buffers, overwritten by the next evaluation step."""
        N = int(n_items) if n_items is not None else int(self.eval_sampler.items.shape[0])
        if batch is None and self._graph_ok(e - s) and e - s > 0 and e <= self.st.n_events:
            sg = self._step_graph(e - s)
            self._fill_static(sg, s, e)
            if ev_base is not None:      # multi-GPU: position of this rank's slice inside the global evaluation batch
                if "ev_base" not in sg.static:
                    sg.static["ev_base"] = torch.zeros(1, dtype=torch.int64, device=self.device)
                sg.static["ev_base"].fill_(int(ev_base))
            if state_batch is not None:
                self._ensure_state_buffers(sg, state_batch["src"].shape[0])
                for k, v in state_batch.items():
                    sg.static["state"][k].copy_(v)
            eg = sg.evals.setdefault((N, state_batch is not None), _StepGraph(sg.static))
            if eg.graph is None:
                if eg.eager_steps < 2:
                    eg.eager_steps += 1
                    return self._eval_body(sg.static, N, state_batch is not None)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                launches0 = _lib.LAUNCHES
                with torch.cuda.graph(g, **self._capture_kw):
                    eg.loss = self._eval_body(sg.static, N, state_batch is not None)
                eg.launches = _lib.LAUNCHES - launches0
                eg.graph = g
            eg.graph.replay()
            _lib.LAUNCHES += eg.launches
            return eg.loss
        b = dict(batch) if batch is not None else self._batch(s, e)
        if state_batch is not None:
            b["state"] = state_batch
        if ev_base is not None:
            b["ev_base"] = int(ev_base)
        return self._eval_body(b, N, state_batch is not None)

    def _ensure_state_buffers(self, sg, Bg):
        if "state" not in sg.static or sg.static["state"]["src"].shape[0] != Bg:
            dev, i32 = self.device, torch.int32
            sg.static["state"] = dict(src=torch.zeros(Bg, dtype=i32, device=dev), dst=torch.zeros(Bg, dtype=i32, device=dev),
                                      ts=torch.zeros(Bg, dtype=torch.float64, device=dev),
                                      eidx=torch.zeros(Bg, dtype=i32, device=dev))

    def _eval_body(self, b, N, with_state):
        D = self.dev_stream
        tgn = self.tgn.eval()
        eng = tgn._get_engine()
        eng.nf = self.nf_full
        B = b["src"].shape[0]
        # evaluation recreates RandomState(2024) per batch: the stream is keyed by position in the (global) batch
        ev = torch.arange(B, dtype=torch.int64, device=self.device)
        if "ev_base" in b:
            ev = ev + b["ev_base"]
        held = b["port_items"] + (self.st.n_users + 1) if "port_items" in b else D.port_items_as_item_ids
        cand = self.eval_sampler.sample(ev, b["port_ptr"], held, N, seed=2024)
        e_s, e_d, e_c = eng.compute_temporal_embeddings(tgn._params(), b["src"], b["dst"], [cand.reshape(-1)], b["ts"],
                                                        b["eidx"], self.tc.n_neighbors, train=False,
                                                        state_batch=b["state"] if with_state else None)
        if self.metrics is not None and N >= 4:
            pos_rank, top, scores, per_event = self.metrics.step(e_s, e_d, e_c, b["dst"], cand, b["day"], b["port_ptr"],
                                                                 b["port_items"] if "port_items" in b else D.port_items)
            return pos_rank, top, cand, scores, per_event
        d = e_s.shape[1]
        scores = torch.empty(B, 1 + N, device=self.device)
        pos_rank = torch.empty(B, dtype=torch.int32, device=self.device)
        top = torch.empty(B, 5, dtype=torch.int32, device=self.device)
        e_s, e_d, e_c = e_s.contiguous(), e_d.contiguous(), e_c.contiguous()
        _lib.call("pfo_eval_score", ptr(e_s), ptr(e_d), ptr(e_c), B, N, d, 5, ptr(scores), ptr(pos_rank), ptr(top))
        return pos_rank, top, cand, scores, None

    def reset_eval_metrics(self):
        if self.metrics is not None:
            self.metrics.reset()

    def eval_summary(self, EVAL="val"):
        """The dictionary reference eval_recommendation returns (evaluation.py:209-258), from the 31 running sums
        the evaluation steps since `reset_eval_metrics` left on the device (in the replicated multi-GPU mode the
        sums of the ranks' user slices are all-reduced first)."""
        if self.metrics is None:         # no price tables on this stream: nothing to summarise
            return {}
        return self.metrics.summary(EVAL, reduce=self._all_ranks_sum)

    def _all_ranks_sum(self, t):
        return t

    def evaluate(self, s, e, bs=None, n_items=None, EVAL="val", max_batches=None):
        """The loop of reference eval_recommendation (evaluation.py:63-207) over interactions [s, e): batches of
        `bs`, the last (short or exactly-ending) batch skipped like the reference does (:68-69); `max_batches` is the
        reference's `is_test_run` stop (:72-74)."""
        bs = int(bs or self.tc.bs)
        self.reset_eval_metrics()
        n_batches = -(-(e - s) // bs)
        for i in range(n_batches):
            a, b_ = s + i * bs, min(e, s + (i + 1) * bs)
            if b_ == e:
                continue
            if max_batches is not None and i >= max_batches:
                break
            self.eval_step(a, b_, n_items=n_items)
        return self.eval_summary(EVAL)

    def split_ranges(self):
        """(train, validation, test) index ranges: the chronological 80/10/10 split of reference utils/data.py:27,48-50
        (the stream is time-sorted, so the three masks are contiguous)."""
        tr, va, _ = self.masks
        if not np.all(np.diff(self.st.timestamps) >= 0):
            # the index ranges below equal the reference's timestamp masks only on a chronological stream; on an
            # unsorted file they would leak validation / test interactions into training
            raise ValueError("the interaction stream is not sorted by timestamp: fit() walks contiguous index ranges")
        n_tr, n_va = int(tr.sum()), int(va.sum())
        return (0, n_tr), (n_tr, n_tr + n_va), (n_tr + n_va, self.st.n_events)

    def fit(self, epochs=1, bs=None, max_batches=None, log=None):
        """The epoch loop of reference main.py:144-443: re-initialise the memory, train over the training split in
        batches of `bs` (the last one may be short), then evaluate the validation and the test split on the full
        graph -- the memory carries over from training into validation into test, as in the reference.  Returns one
        dictionary per epoch: the mean training loss plus the reference's 30 `valid_*` and 30 `test_*` keys."""
        bs = int(bs or self.tc.bs)
        (t0, t1), (v0, v1), (e0, e1) = self.split_ranges()
        history = []
        self._check_fit(bs)
        for epoch in range(int(epochs)):
            self.epoch = epoch
            if self.tgn.use_memory:
                self.tgn.memory.__init_memory__()                    # main.py:152-153
            losses = []
            for bi in range(-(-(t1 - t0) // bs)):
                if max_batches is not None and bi >= max_batches:    # main.py:162-164 (--test_run)
                    break
                s, e = t0 + bi * bs, min(t1, t0 + (bi + 1) * bs)
                losses.append(self.train_step(s, e).clone())         # the graph's output buffer is reused next step
            out = {"epoch": epoch, "loss": float(torch.stack(losses).mean().item()) if losses else float("nan")}
            out.update(self.evaluate(v0, v1, bs=bs, EVAL="valid", max_batches=max_batches))   # main.py:405-418
            out.update(self.evaluate(e0, e1, bs=bs, EVAL="test", max_batches=max_batches))    # main.py:427-440
            history.append(out)
            if log is not None:
                log(out)
        return history

    def _check_fit(self, bs):
        pass

    @staticmethod
    def recall_ndcg(pos_rank, ks=(1, 3, 5)):
        """Recall@k / NDCG@k with a single relevant item (reference evaluation.py:11-21,141-144)."""
        r = pos_rank.to(torch.float32)
        out = {}
        for k in ks:
            hit = (r < k).to(torch.float32)
            out[f"recall_{k}"] = hit.mean()
            out[f"ndcg_{k}"] = (hit / torch.log2(r + 2.0)).mean()
        return out


def replica_slice(s, e, rank, world):
    """This rank's share [ls, le) of the global batch [s, e): consecutive slices, the first (n mod world) ranks take
    one interaction more (the short last batch of an epoch, main.py:180, rarely divides)."""
    n = e - s
    if n < world:
        raise ValueError(f"global batch of {n} interactions cannot be split over {world} ranks")
    base, rem = divmod(n, world)
    ls = s + rank * base + min(rank, rem)
    return ls, ls + base + (1 if rank < rem else 0)


def allreduce_sum_(flat, group=None):
    """In-place sum of the flat gradient bucket over the ranks (the loss is pre-scaled by 1 / world)."""
    import torch.distributed as dist
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class ReplicatedTrainer(PfoTrainer):
    """Data-parallel training over `world` GPUs with REPLICATED state (one process per GPU).

    Every rank holds the whole node memory, pending-message table and adjacency (config 4 of BASELINE.json --
    10M users, 1B events -- is ~45 GB of state: it fits each 180 GB B200), embeds ITS slice of the global batch
    (sampling, attention, BPR, backward are embarrassingly parallel over interactions) and advances the state with
    the WHOLE global batch, which it already has on the device: the lazily updated memory rows of all positives
    are part of its node table, so persist + last-wins message store run identically on every replica and the
    replicas never diverge.  The only exchange per step is one NCCL all-reduce of the 133 k-float gradient bucket,
    launched right after the step's CUDA graph (sampling .. backward), followed by the fused Adam.  Same semantics as the reference on the global batch (embeddings from
    the pre-batch state, then one state update), same numbers as the 1-GPU step on that batch up to the order of
    the gradient sum (tools/check_replicated.py).  The node-sharded alternative (all-to-all routing, for state
    beyond one GPU) is pfotgnrec_b200/dist.py."""

    def __init__(self, st, tc, device, rank, world, group=None, nccl_in_graph=False):
        super().__init__(st, tc, device)
        if tc.model in ("dyrep",):
            raise NotImplementedError("dyrep messages carry embeddings of the other ranks' interactions")
        self.rank, self.world, self.group, self.nccl_in_graph = int(rank), int(world), group, bool(nccl_in_graph)
        self._loss_scale = 1.0 / self.world
        # nccl_in_graph=False: the graph holds sampling .. backward, the all-reduce and Adam are launched after it
        self._graph_tail_eager = not self.nccl_in_graph
        self.tgn._get_engine().seed = tc.seed + 7919 * self.rank        # decorrelate the dropout streams of the replicas
        self.params = [p for p in self.tgn.parameters() if p.requires_grad]
        self._own_bucket = self.gflat is None
        if self._own_bucket:         # operands derived by torch ops (graph_sum): autograd accumulates into views of a bucket
            sizes = [p.numel() for p in self.params]
            self.gflat = torch.zeros(sum(sizes), device=self.device)
            for p, g in zip(self.params, self.gflat.split(sizes)):
                p.grad = g.view_as(p)

    def _zero_grads(self):
        if self._own_bucket:
            self.gflat.zero_()

    def _check_fit(self, bs):
        if bs < self.world:
            raise ValueError(f"fit: batch size {bs} is smaller than the number of ranks ({self.world})")

    def _all_ranks_sum(self, t):
        return allreduce_sum_(t, self.group) if self.world > 1 else t

    def _reduce_grads(self):
        if self.world > 1:
            allreduce_sum_(self.gflat, self.group)

    def _step_graph(self, B):
        sg = super()._step_graph(B)
        self._ensure_state_buffers(sg, B * self.world)
        return sg

    def eval_step(self, s, e, n_items=None, batch=None, state_batch=None):
        """Global evaluation batch [s, e): this rank scores its slice of the users against all candidates, every
        rank advances the state with the whole batch (returns this rank's slice of the results)."""
        ls, le = replica_slice(s, e, self.rank, self.world)
        return super().eval_step(ls, le, n_items=n_items, state_batch=self._state_batch(s, e), ev_base=ls - s)

    def _state_batch(self, s, e):
        D = self.dev_stream
        return dict(src=D.src[s:e], dst=D.dst[s:e], ts=D.ts[s:e], eidx=D.eidx[s:e])

    def train_step(self, s, e, batch=None):
        """Global batch [s, e): this rank embeds its slice, every rank advances the state with all of it."""
        ls, le = replica_slice(s, e, self.rank, self.world)
        if (e - s) % self.world != 0:
            # ragged tail batch: the loss is the mean over the GLOBAL batch, so this rank's mean is weighted by its
            # share; launched kernel by kernel (the captured graphs bake the even 1 / world weight)
            b = self._batch(ls, le)
            b["state"] = self._state_batch(s, e)
            keep, self._loss_scale = self._loss_scale, (le - ls) / float(e - s)
            try:
                return self._step_body(b)
            finally:
                self._loss_scale = keep
        if self._graph_ok(le - ls) and e <= self.st.n_events:
            sg = self._step_graph(le - ls)
            self._fill_static(sg, ls, le)
            for k, v in self._state_batch(s, e).items():
                sg.static["state"][k].copy_(v)
            return self._run_graphed(sg)
        b = self._batch(ls, le)
        b["state"] = self._state_batch(s, e)
        return self._step_body(b)

    def make_host_batches(self, start, count, bs):
        """Pinned host copies of `count` consecutive GLOBAL batches of bs * world interactions: this rank's slice
        (all columns) plus the four columns of the whole batch that advance the state."""
        st, out = self.st, []
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        for i in range(count):
            s = start + i * bs * self.world
            e = s + bs * self.world
            ls, le = replica_slice(s, e, self.rank, self.world)
            hb = super().make_host_batches(ls, 1, bs)[0]
            hb["state"] = dict(src=pin(st.sources[s:e].astype(np.int32)), dst=pin(st.destinations[s:e].astype(np.int32)),
                               ts=pin(st.timestamps[s:e]), eidx=pin(st.edge_idxs[s:e].astype(np.int32)))
            hb["nbytes"] += sum(v.numel() * v.element_size() for v in hb["state"].values())
            out.append(hb)
        return out

    def train_step_host(self, hb):
        B = hb["src"].shape[0]
        if self._graph_ok(B):
            sg = self._step_graph(B)
            self._copy_host_batch(sg, {k: v for k, v in hb.items() if k != "state"})
            for k, v in hb["state"].items():
                sg.static["state"][k].copy_(v, non_blocking=True)
            return self._run_graphed(sg)
        b = {k: v.to(self.device, non_blocking=True) for k, v in hb.items() if k not in ("nbytes", "state", "_packed")}
        b["state"] = {k: v.to(self.device, non_blocking=True) for k, v in hb["state"].items()}
        return self._step_body(b)

"""Step engine: dense TGN state on the device + the kernel sequence of one batch.

Host-side orchestration of the path reference model/tgn.py:102-327 walks
(get_updated_memory -> compute_embedding -> update_memory -> get_raw_messages -> store), on
the restatements of SURVEY.md section 7: dense pending-message table, lazy memory update on
the unique touched nodes only, one cell evaluation for both the functional and the persisted
view, CSR adjacency.  All arithmetic runs in libpfo_b200.so; torch is used for buffers,
streams and the autograd plumbing (`TGNStepFunction`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import ptr

CELL_GRU, CELL_RNN, CELL_NONE = 0, 1, 2


@dataclass
class ModelConfig:
    d: int                      # node-feature / memory / time-encoding dimension
    n_edge_feat: int
    n_layers: int = 1
    n_heads: int = 2
    use_memory: bool = True
    updater: str = "gru"        # "gru" | "rnn"
    embedding: str = "graph_attention"   # "graph_attention" | "graph_sum" | "time" | "identity"
    dyrep: bool = False
    dst_emb_in_msg: bool = False
    src_emb_in_msg: bool = False
    message_fn: str = "identity"         # "identity" | "mlp" (reference modules/message_function.py:13-33)
    msg_dim: int = 100                   # output width of the MLP message function
    aggregator: str = "last"             # "last" | "mean" (reference modules/message_aggregator.py:38-81)
    shift: tuple = (0.0, 1.0, 0.0, 1.0)  # mean/std time shift src, dst
    dropout: float = 0.0
    gemm_mode: str = "fp32"     # "fp32" (3xTF32 tcgen05, 1e-5) | "tf32" (tcgen05, 2e-2) | "bf16" | "simt" (FFMA)
    # memory updater GEMMs: "split" = torch's two (W_ih . message, W_hh . memory); "merged" = ONE contraction over the
    # operand row [message | memory] with a block weight.  Measured at bs 8192 (profiles/r2_bench_merged_cell.json): the
    # merged forward GEMM saves 15 us, but the merged weight gradient [4d x (raw + d)] costs 118 us against 74 us for the
    # two separate ones (the wgrad kernel's slab decomposition degrades at N = 256), and the bf16 kernel does not take
    # the shape -- so "split" stays the default and "merged" is an opt-in covered by its own parity test.
    cell_gemm: str = "split"

    @property
    def E(self):                # attention embed dim = d + time dim
        return 2 * self.d

    @property
    def Ek(self):               # key / value input dim
        return 2 * self.d + self.n_edge_feat

    @property
    def ekp(self):              # per-head row of QK / XB: [h | e | te | psum | valid | one | 0..], 16-byte multiple
        return (self.Ek + 3 + 3) // 4 * 4

    @property
    def ldcat(self):            # operand row of the merge GEMM: [XB (H * ekp) | h_query (d)]
        return self.n_heads * self.ekp + self.d

    @property
    def raw(self):              # raw message width
        return 3 * self.d + self.n_edge_feat

    @property
    def rawp(self):
        return (self.raw + 3) // 4 * 4

    @property
    def graph(self):            # embeddings that walk the sampled temporal neighbourhood
        return self.embedding in ("graph_attention", "graph_sum")

    @property
    def ekp1(self):             # single-head row of the neighbour kernels (graph_sum)
        return self.ekp

    @property
    def cell_in(self):          # input width of the memory updater
        return self.raw if self.message_fn == "identity" else self.msg_dim

    @property
    def mlp_hidden(self):       # nn.Linear(raw, raw // 2) of the MLP message function
        return self.raw // 2

    @property
    def cell(self):
        if not self.use_memory:
            return CELL_NONE
        return CELL_GRU if self.updater == "gru" else CELL_RNN

    @property
    def gates(self):
        return 3 if self.updater == "gru" else 1


class _FoldAttention(torch.autograd.Function):
    """(Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb) -> (Wqk, cqk, Wc1T) and its adjoint, two kernel launches each way."""

    @staticmethod
    def forward(ctx, cfg, Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb):
        d, F, H, ekp = cfg.d, cfg.n_edge_feat, cfg.n_heads, cfg.ekp
        dev = Wq.device
        w = [t.detach().contiguous() for t in (Wq, Wk, Wv, b_in, Wo, bo, W1, b1, tb)]
        ws = torch.empty(int(_lib.query("pfo_fold_attention_workspace_doubles", d, F, H)), dtype=torch.float64, device=dev)
        Wqk = torch.empty(H * ekp, d, device=dev)
        cqk = torch.empty(H * ekp, device=dev)
        Wc1T = torch.empty(H * ekp + d, d, device=dev)
        _lib.call("pfo_fold_attention_fwd", *[ptr(t) for t in w], d, F, H, ekp, ptr(ws), ptr(Wqk), ptr(cqk), ptr(Wc1T))
        ctx.cfg, ctx.w, ctx.ws = cfg, w, ws
        return Wqk, cqk, Wc1T

    @staticmethod
    def backward(ctx, gWqk, gcqk, gWc1T):
        cfg, w, ws = ctx.cfg, ctx.w, ctx.ws
        d, F, H, ekp = cfg.d, cfg.n_edge_feat, cfg.n_heads, cfg.ekp
        g_in = [gWqk.contiguous(), gcqk.contiguous(), gWc1T.contiguous()]
        g_out = [torch.empty_like(t) for t in w]
        _lib.call("pfo_fold_attention_bwd", *[ptr(t) for t in w], d, F, H, ekp, ptr(ws), *[ptr(t) for t in g_in],
                  *[ptr(t) for t in g_out])
        return (None,) + tuple(g_out)


class TGNState:
    """memory / last_update / pending messages of every node, resident in HBM.

    Dense restatement of reference modules/memory.py:8-75 (`Memory.memory`, `.last_update`,
    `.messages`)."""

    def __init__(self, n_nodes: int, cfg: ModelConfig, device):
        self.n_nodes, self.cfg, self.device = n_nodes, cfg, torch.device(device)
        N, d = n_nodes, cfg.d
        dev = self.device
        self.memory = torch.zeros(N, d, device=dev)
        self.last_update = torch.zeros(N, device=dev)
        self.pend_msg = torch.zeros(N, cfg.rawp, device=dev)
        self.pend_ts = torch.zeros(N, device=dev)
        self.pend_valid = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.last_pos = torch.full((N,), -1, dtype=torch.int32, device=dev)
        self.bitmap = torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev)
        self.slot_of_node = torch.zeros(N, dtype=torch.int32, device=dev)
        self.compact_ws = torch.zeros(int(_lib.query("pfo_compact_workspace_ints", N)) if dev.type == "cuda" else 1,
                                      dtype=torch.int32, device=dev)
        self.n_unique = torch.zeros(1, dtype=torch.int32, device=dev)

    def reset(self):
        """reference modules/memory.py:23-33 (__init_memory__)."""
        self.memory.zero_()
        self.last_update.zero_()
        self.pend_valid.zero_()
        self.pend_ts.zero_()
        self.pend_msg.zero_()

    def backup(self):
        return tuple(t.clone() for t in (self.memory, self.last_update, self.pend_msg, self.pend_ts, self.pend_valid))

    def restore(self, b):
        for dst, src in zip((self.memory, self.last_update, self.pend_msg, self.pend_ts, self.pend_valid), b):
            dst.copy_(src)


_LINEAR_PASSES = {"fp32": 3, "tf32": 1}


def _linear(cfg, A, lda, a_idx, W, ldw, wt, bias, C, ldc, M, N, K, *, m_dev=None, alpha=1.0, act=0,
            row_zero=None, relu_gate=None, ld_gate=0, accumulate=0, brs=None, ld_brs=0):
    """C = epi(alpha * (A W^T + bias)).  gemm_mode: "fp32" = 3xTF32 on tcgen05 (1e-5 contract),
    "tf32" = single-pass TF32 on tcgen05, "bf16" = bf16 tcgen05, "simt" = FFMA."""
    mode = cfg.gemm_mode
    if mode in _LINEAR_PASSES:
        _lib.call("pfo_linear_tf32", A, lda, a_idx, W, ldw, int(wt), bias, brs, ld_brs, C, ldc, M, m_dev, N, K,
                  float(alpha), int(act), row_zero, relu_gate, ld_gate, int(accumulate), _LINEAR_PASSES[mode])
        return
    name = "pfo_linear_f32" if mode == "simt" else "pfo_linear_bf16"
    _lib.call(name, A, lda, a_idx, W, ldw, int(wt), bias, brs, ld_brs, C, ldc, M, m_dev, N, K,
              float(alpha), int(act), row_zero, relu_gate, ld_gate, int(accumulate))


class _Workspace:
    """Scratch for the two-stage deterministic reductions."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, n_floats):
        if self.buf is None or self.buf.numel() < n_floats:
            self.buf = torch.empty(max(int(n_floats), 1 << 20), device=self.device)
        return self.buf


def _wgrad(cfg, ws, G, ldg, A, lda, a_idx, M, N, K, dW, lddw, db, *, m_dev=None, accumulate=0):
    """dW = G^T A (+ db): tcgen05 in the "fp32" (3xTF32) / "tf32" modes, FFMA otherwise."""
    wb = 1 if db is not None else 0
    if cfg.gemm_mode in _LINEAR_PASSES:
        buf = ws.get(_lib.query("pfo_wgrad_tf32_workspace_floats", M, N, K, wb))
        _lib.call("pfo_wgrad_tf32", G, ldg, A, lda, a_idx, M, m_dev, N, K, dW, lddw, db, int(accumulate), ptr(buf),
                  _LINEAR_PASSES[cfg.gemm_mode])
        return
    buf = ws.get(_lib.query("pfo_wgrad_workspace_floats", M, N, K, wb))
    _lib.call("pfo_wgrad_f32", G, ldg, A, lda, a_idx, M, m_dev, N, K, dW, lddw, db, int(accumulate), ptr(buf))


F4 = 4  # sizeof(float)


class _LayerTape:
    """Everything one attention-layer invocation saves for its backward."""
    __slots__ = ("layer", "M", "n", "Tq", "qidx", "T", "idx", "eidx", "dt", "CAT", "QK", "P",
                 "invalid", "H1", "out_rows", "child_q", "child_n", "dTq", "dT", "step")


class TGNEngine:
    """One model instance bound to a state, node/edge feature tables and a neighbour finder."""

    def __init__(self, cfg: ModelConfig, state: Optional[TGNState], node_feat: torch.Tensor,
                 edge_feat: torch.Tensor, neighbor_finder):
        # the limits the kernels are specialised for, named here: past this point a violation would only surface as
        # cudaErrorInvalidValue from an entry point
        if cfg.d % 32 != 0 or not 32 <= cfg.d <= 128:
            raise ValueError(f"memory / feature dimension d={cfg.d}: the kernels are specialised for d in {{32, 64, 96, 128}}")
        if cfg.embedding == "graph_attention" and (cfg.n_heads < 1 or cfg.n_heads > 4 or cfg.E % cfg.n_heads != 0):
            raise ValueError(f"n_heads={cfg.n_heads}: the attention kernels take 1..4 heads that divide 2 * d = {cfg.E}")
        self.cfg, self.state = cfg, state
        self.node_feat = node_feat.contiguous()
        self.edge_feat = edge_feat.contiguous()
        self.nf = neighbor_finder
        self.device = _lib.use_device(node_feat.device)
        self.n_nodes = node_feat.shape[0]
        self.ws = _Workspace(self.device)
        # graph_sum sums over ALL n sampled slots, padded ones included (they carry node 0's features, the edge
        # feature row 0 and te(t - 0), embedding_module.py:205-208): node 0 then needs a row in the node table
        self.skip_zero = 0 if cfg.embedding == "graph_sum" else 1
        self._slot_cache = {}
        # optional flat gradient buffer in the order of `param_names()` (set by the trainers): the step's backward then
        # writes every parameter gradient straight into it and hands autograd nothing to accumulate
        self.grad_sink = None
        self.step_id = 0
        self.seed = 0
        # device-resident batch counter keying the dropout stream (bumped on the stream each batch, so a
        # captured CUDA graph of the step draws fresh masks on every replay)
        self.step_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        # training steps of the trainers (a backward pass always follows the forward pass): persist + message store run
        # on the side stream beside the loss and the backward pass -- nothing there reads the state they write -- and are
        # joined at the end of the backward pass (the sharded engine joins before its gradient exchange).  Off by
        # default: a caller of the drop-in TGN may read the memory between its forward and backward calls on the main
        # stream; trainer.PfoTrainer turns it on (measured: 2 us of a 0.82 ms step, profiles/r2_knob_sweeps.txt).
        self.overlap_store = False
        self._side_pending = False
        if state is None:        # memory-less models still need the compaction scratch
            self.state = TGNState(self.n_nodes, ModelConfig(d=cfg.d, n_edge_feat=cfg.n_edge_feat), self.device)

    # ------------------------------------------------------------------ parameter packing
    def param_names(self):
        c = self.cfg
        names = ["time_encoder.w.weight", "time_encoder.w.bias"]
        if c.use_memory:
            pre = "memory_updater.memory_updater."
            names += [pre + "weight_ih", pre + "weight_hh", pre + "bias_ih", pre + "bias_hh"]
            if c.message_fn == "mlp":
                names += ["message_function.mlp.0.weight", "message_function.mlp.0.bias",
                          "message_function.mlp.2.weight", "message_function.mlp.2.bias"]
        if c.embedding == "graph_attention":
            for l in range(c.n_layers):
                a = f"embedding_module.attention_models.{l}."
                names += [a + "multi_head_target.q_proj_weight", a + "multi_head_target.k_proj_weight",
                          a + "multi_head_target.v_proj_weight", a + "multi_head_target.in_proj_bias",
                          a + "multi_head_target.out_proj.weight", a + "multi_head_target.out_proj.bias",
                          a + "merger.fc1.weight", a + "merger.fc1.bias", a + "merger.fc2.weight",
                          a + "merger.fc2.bias"]
        elif c.embedding == "graph_sum":
            for l in range(c.n_layers):
                names += [f"embedding_module.linear_1.{l}.weight", f"embedding_module.linear_1.{l}.bias",
                          f"embedding_module.linear_2.{l}.weight", f"embedding_module.linear_2.{l}.bias"]
        elif c.embedding == "time":
            names += ["embedding_module.embedding_layer.weight", "embedding_module.embedding_layer.bias"]
        return names

    def flat_layout_ok(self):
        """True when every operand of the step is one reference parameter as it is (`_pack` adds no arithmetic), i.e.
        the flat buffers of the trainer can follow `param_names()`.  graph_sum derives its operands with torch ops."""
        return self.cfg.embedding != "graph_sum"

    def _pack(self, params):
        """Derived, kernel-friendly tensors built with torch autograd ops on the reference-named
        parameters (tiny; gradients flow back through them to the parameters)."""
        c = self.cfg
        d, E, Ek, H = c.d, c.E, c.Ek, c.n_heads
        hd = E // H
        tw = params["time_encoder.w.weight"].reshape(d)
        tb = params["time_encoder.w.bias"]
        flat = [tw.contiguous(), tb.contiguous()]
        if c.use_memory:
            pre = "memory_updater.memory_updater."
            flat += [params[pre + "weight_ih"].contiguous(), params[pre + "weight_hh"].contiguous(),
                     params[pre + "bias_ih"].contiguous(), params[pre + "bias_hh"].contiguous()]
            if c.message_fn == "mlp":
                flat += [params["message_function.mlp.%s" % k].contiguous() for k in ("0.weight", "0.bias", "2.weight", "2.bias")]
        if c.embedding == "graph_attention":
            flat += self._layer_params(params)
        elif c.embedding == "graph_sum":
            # GraphSumEmbedding.aggregate (embedding_module.py:205-219) in the operand order of the neighbour kernel:
            #   sum_j linear_1([h_j | te_j | e_j]) = n * (W1p . mean_j [h_j | e_j | te_j] + b1)   (columns permuted)
            #   linear_2([sum | h_q | te(0)])     = W2a . [sum | h_q] + (b2 + W2[:, 2d:] cos(tb))   (te(0) = cos(b))
            # built with torch ops on the parameters, so autograd carries the gradients back to them
            F_ = c.n_edge_feat
            for l in range(c.n_layers):
                W1 = params[f"embedding_module.linear_1.{l}.weight"]
                W2 = params[f"embedding_module.linear_2.{l}.weight"]
                flat += [torch.cat([W1[:, :d], W1[:, 2 * d:2 * d + F_], W1[:, d:2 * d]], dim=1).contiguous(),
                         params[f"embedding_module.linear_1.{l}.bias"].contiguous(),
                         W2[:, :2 * d].contiguous(),
                         params[f"embedding_module.linear_2.{l}.bias"] + W2[:, 2 * d:] @ torch.cos(tb)]
        elif c.embedding == "time":
            flat += [params["embedding_module.embedding_layer.weight"].reshape(d).contiguous(),
                     params["embedding_module.embedding_layer.bias"].contiguous()]
        return flat

    def _layer_params(self, params):
        """The reference's ten tensors per attention layer, in the order the fold kernels take them
        (model/temporal_attention.py:26-32 nn.MultiheadAttention + MergeLayer utils/utils.py:7-8)."""
        out = []
        for l in range(self.cfg.n_layers):
            a = f"embedding_module.attention_models.{l}."
            m = a + "multi_head_target."
            out += [params[k].contiguous() for k in (
                m + "q_proj_weight", m + "k_proj_weight", m + "v_proj_weight", m + "in_proj_bias",
                m + "out_proj.weight", m + "out_proj.bias", a + "merger.fc1.weight", a + "merger.fc1.bias",
                a + "merger.fc2.weight", a + "merger.fc2.bias")]
        return out

    def fold_layers(self, raw, tb):
        """Per layer, the five GEMM operands the kernels consume, folded from the reference's ten tensors
        (pfo_fold_attention_fwd, csrc/fold_kernels.cu; fp64 arithmetic on the weights):

          Wqk  [H*ekp, d], cqk [H*ekp]   qk_h = (scale Wk_h^T Wq_h[:, :d]) h_q + scale Wk_h^T (Wq_h[:, d:] te(0) + bq_h)
          Wc1T [H*ekp + d, d]            fc1([out_proj(attn) | h_q]) as ONE contraction over [XB | h_q]:
                                         rows of head h = (W1a Wo_h [Wv_h | bv_h])^T, then the `valid` row = W1a bo
                                         (out-proj bias, valid rows only) and the `one` row = b1; last d rows = W1b^T
          W2, b2                         fc2
        K/V projections, the query projection, the out-projection and fc1 never materialise per query
        (model/temporal_attention.py:52-90, utils/utils.py:14-17).  The fold only touches weights, so it runs on
        a side stream (a parallel branch of the captured graph) while the batch is sampled; `join_fold` is the
        join point.  Returns (folded operands per layer, fp64 workspaces kept for the adjoint)."""
        c, dev = self.cfg, self.device
        d, F, H, ekp = c.d, c.n_edge_feat, c.n_heads, c.ekp
        n_ws = int(_lib.query("pfo_fold_attention_workspace_doubles", d, F, H))
        folded, wss = [], []
        for l in range(c.n_layers):
            w = raw[l]
            wss.append(torch.empty(n_ws, dtype=torch.float64, device=dev))
            folded.append([torch.empty(H * ekp, d, device=dev), torch.empty(H * ekp, device=dev),
                           torch.empty(H * ekp + d, d, device=dev), w[8], w[9]])
        cur = torch.cuda.current_stream(dev)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for l in range(c.n_layers):
                w, f = raw[l], folded[l]
                _lib.call("pfo_fold_attention_fwd", *[ptr(t) for t in w[:8]], ptr(tb), d, F, H, ekp, ptr(wss[l]),
                          ptr(f[0]), ptr(f[1]), ptr(f[2]))
        return folded, wss

    def join_side(self):
        torch.cuda.current_stream(self.device).wait_stream(self.side)

    def unfold_layer_grads(self, raw, tb, wss, g_folded, g_raw, g_tb_fold):
        """Adjoint of `fold_layers` on the side stream: d/d(Wqk, cqk, Wc1T) -> d/d(the ten reference tensors)."""
        c, dev = self.cfg, self.device
        d, F, H, ekp = c.d, c.n_edge_feat, c.n_heads, c.ekp
        self.side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.side):
            for l in range(c.n_layers):
                w, gf, gr = raw[l], g_folded[l], g_raw[l]
                _lib.call("pfo_fold_attention_bwd", *[ptr(t) for t in w[:8]], ptr(tb), d, F, H, ekp, ptr(wss[l]),
                          ptr(gf[0]), ptr(gf[1]), ptr(gf[2]), *[ptr(t) for t in gr[:8]], ptr(g_tb_fold[l]))

    # ------------------------------------------------------------------ public step
    def compute_temporal_embeddings(self, params, src, dst, extra_groups, ts, eidx, n_neighbors,
                                    train=True, update_state=True, state_batch=None, packed=False):
        """src/dst int32[B], extra_groups list of int32[k*B] (interaction-major), ts float64[B],
        eidx int32[B] -- all device tensors.  Returns [emb_src, emb_dst, *emb_extra] ([.,d] fp32,
        differentiable w.r.t. `params` when grad mode is on) and advances memory / messages.

        state_batch (optional, dict(src, dst, ts, eidx)): the interactions whose memory / messages are advanced, when
        they are not the ones embedded -- the replicated data-parallel mode embeds this rank's slice of the global
        batch and advances the (replicated) state with the WHOLE global batch, so every replica stays identical
        without exchanging rows."""
        groups = [src, dst] + list(extra_groups)
        B = src.shape[0]
        if self.cfg.graph and int(n_neighbors) > 32:
            raise ValueError(f"n_neighbors={n_neighbors}: the neighbour kernels hold one sampled slot per lane (at most 32)")
        q_nodes = torch.cat(groups)
        q_ts = torch.cat([ts if g.shape[0] == B else ts.repeat_interleave(g.shape[0] // B) for g in groups])
        flat = self._pack(params)
        batch = dict(src=src, dst=dst, ts=ts, eidx=eidx, q_nodes=q_nodes, q_ts=q_ts, n=int(n_neighbors),
                     B=B, train=bool(train), update_state=bool(update_state), q_ids=self._query_ids(groups, B),
                     grad=torch.is_grad_enabled())      # inside Function.forward grad mode is always off: ask here
        if state_batch is not None:
            if self.cfg.dst_emb_in_msg or self.cfg.src_emb_in_msg:
                raise NotImplementedError("messages that carry embeddings (dyrep) need the embedded batch as state batch")
            batch["state"] = dict(src=state_batch["src"], dst=state_batch["dst"], ts=state_batch["ts"],
                                  eidx=state_batch["eidx"], B=int(state_batch["src"].shape[0]))
        emb = TGNStepFunction.apply(self, batch, *flat)
        if packed:          # one [Q, d] tensor, rows in group order: a loss over row blocks hands ONE gradient back
            return emb      # (splitting costs autograd a zero fill and a concatenation of the pieces per step)
        return list(torch.split(emb, [g.shape[0] for g in groups]))

    # ------------------------------------------------------------------ forward internals
    def _query_ids(self, groups, B):
        """Ids of the queries in the un-sharded query list (only the node-sharded engine needs them: they key the
        uniform neighbour stream, so that the draws do not depend on the number of ranks)."""
        return None

    def _sample_tree(self, nodes, ts, layer, n, q_ids=None):
        """Neighbour sampling in the call order of the reference recursion
        (modules/embedding_module.py:115-145): inner queries, own call, inner neighbours."""
        if layer == 0:
            return None
        M = nodes.shape[0]
        child_q = self._sample_tree(nodes, ts, layer - 1, n, q_ids)
        if q_ids is not None:
            nbr, eidx, _etime, dt = self.nf.sample(nodes, ts, n, q_ids=q_ids)
        else:
            nbr, eidx, _etime, dt = self.nf.sample(nodes, ts, n)
        n_eff = nbr.shape[1]
        child_n = None
        if layer > 1:
            # a neighbour's id = its position in the flattened [M, n] list of the un-sharded call
            ids_n = None if q_ids is None else (q_ids.view(-1, 1) * n_eff
                                                + torch.arange(n_eff, dtype=torch.int32, device=nodes.device)).reshape(-1)
            child_n = self._sample_tree(nbr.reshape(-1), ts.repeat_interleave(n_eff), layer - 1, n, ids_n)
        return dict(layer=layer, M=M, nodes=nodes, nbr=nbr, eidx=eidx, dt=dt, child_q=child_q, child_n=child_n)

    def _collect_level0(self, tree, out):
        if tree["layer"] == 1:
            out.append(tree["nodes"])
            out.append(tree["nbr"].reshape(-1))
        else:
            self._collect_level0(tree["child_q"], out)
            self._collect_level0(tree["child_n"], out)

    def _unique_nodes(self, id_lists, scratch=None, n_nodes=None, skip_zero=None):
        """Ascending unique ids of `id_lists` (bitmap marks -> scan -> compact).  `scratch` holds the bitmap, the
        slot-of-node map, the scan workspace and the device-side count (default: the state's own, over its nodes)."""
        st = self.state if scratch is None else scratch
        n_nodes = self.n_nodes if n_nodes is None else n_nodes
        skip = self.skip_zero if skip_zero is None else skip_zero
        total = 0
        for ids in id_lists:
            _lib.call("pfo_mark_nodes", ptr(ids), ids.numel(), skip, ptr(st.bitmap))
            total += ids.numel()
        u_max = min(total, n_nodes)
        uniq = torch.zeros(u_max, dtype=torch.int32, device=self.device)
        _lib.call("pfo_compact_nodes", ptr(st.bitmap), n_nodes, ptr(st.compact_ws), ptr(uniq),
                  ptr(st.slot_of_node), ptr(st.n_unique))
        return uniq, u_max

    # ------------------------------------------------------------------ overridable stages
    # (pfotgnrec_b200/dist.py replaces these three with their node-sharded, all-to-all versions)
    def node_table(self, id_lists, cellW, mlpW=None):
        """Unique touched nodes of the batch and their feature rows: Hnew = lazily updated memory
        (GRU/RNN applied to the pending message, through the MLP message function when configured:
        tgn.py:342-354 aggregate -> compute_message -> updater), H0 = Hnew + node features, lu_u = last_update'."""
        uniq, u_max = self._unique_nodes(id_lists)
        n_uniq = self.state.n_unique.clone()
        self.slot_map = self.state.slot_of_node
        return self._cell_table(uniq, n_uniq, u_max, cellW, mlpW)

    def _cell_table(self, uniq, n_uniq, u_max, cellW, mlpW=None):
        """The memory updater on the rows `uniq` of this engine's state (modules/memory_updater.py:35-53).
        cfg.cell_gemm == "merged": ONE contraction over the operand row [cell input | memory] with the block weight of
        pfo_pack_cell (the operand is read once, one GEMM launch instead of two each way); "split": torch's two GEMMs."""
        c, st, dev = self.cfg, self.state, self.device
        d = c.d
        G = c.gates * d
        H0 = torch.empty(u_max, d, device=dev)
        tab = dict(uniq=uniq, u_max=u_max, n_uniq=n_uniq, H0=H0, Hnew=None, lu_u=None, HG=None, XG=None, valid_u=None,
                   GI=None, GH=None, M1=None, X2=None, XH=None, Wc=None, merged=False)
        if not c.use_memory:
            _lib.call("pfo_cell_forward", ptr(uniq), ptr(n_uniq), u_max, d, c.cell, 0, None, None, G, None, None,
                      ptr(self.node_feat), None, ptr(H0))
            return tab
        merged = c.cell_gemm == "merged"
        mlp = c.message_fn == "mlp"
        HG = torch.empty(u_max, d, device=dev)
        valid_u = torch.empty(u_max, dtype=torch.uint8, device=dev)
        lu_u = torch.empty(u_max, device=dev)
        hid, md = c.mlp_hidden, c.msg_dim
        ld1, ld2 = (hid + 3) // 4 * 4, (md + 3) // 4 * 4
        kx, kxp = (md, ld2) if mlp else (c.raw, c.rawp)     # width of the cell input and of its slot in the operand row
        ldx = kxp + d
        XH = None
        if merged:                                          # operand rows [cell input (kxp) | memory (d)]
            XH = (torch.zeros if mlp else torch.empty)(u_max, ldx, device=dev)
        XG = torch.empty(u_max, c.rawp, device=dev) if (mlp or not merged) else XH
        _lib.call("pfo_gather_state", ptr(uniq), ptr(n_uniq), u_max, d, c.raw, ptr(st.memory), ptr(st.pend_msg),
                  c.rawp, ptr(st.pend_valid), ptr(st.pend_ts), ptr(st.last_update),
                  ptr(HG), ptr(XG), c.rawp if XG is not XH else ldx,
                  (XH.data_ptr() + kxp * F4) if merged else None, ldx, ptr(valid_u), ptr(lu_u))
        W_ih, W_hh, b_ih, b_hh = cellW
        X, ldxin = XG, c.rawp
        M1 = X2 = None
        if mlp:          # Linear(raw, raw // 2) -> ReLU -> Linear(raw // 2, msg_dim)
            W1, b1, W2, b2 = mlpW
            M1 = torch.empty(u_max, ld1, device=dev)
            _linear(c, ptr(XG), c.rawp, None, ptr(W1), c.raw, 0, ptr(b1), ptr(M1), ld1, u_max, hid, c.raw,
                    m_dev=ptr(n_uniq), act=1)
            if merged:
                X2, ldxin = XH, ldx                         # the message lands in its slot of the operand row
            else:
                X2, ldxin = torch.empty(u_max, ld2, device=dev), ld2
            _linear(c, ptr(M1), ld1, None, ptr(W2), hid, 0, ptr(b2), ptr(X2), ldxin, u_max, md, hid, m_dev=ptr(n_uniq))
            X = X2
        Hnew = torch.empty(u_max, d, device=dev)
        GI = GH = Wc = None
        if merged:
            Gm = 4 * d if c.cell == CELL_GRU else d
            Wc = torch.empty(Gm, ldx, device=dev)
            bc = torch.empty(Gm, device=dev)
            _lib.call("pfo_pack_cell", ptr(W_ih), ptr(W_hh), ptr(b_ih), ptr(b_hh), d, kx, kxp, c.cell, ptr(Wc), ptr(bc))
            GI = torch.empty(u_max, Gm, device=dev)
            _linear(c, ptr(XH), ldx, None, ptr(Wc), ldx, 0, ptr(bc), ptr(GI), Gm, u_max, Gm, ldx, m_dev=ptr(n_uniq))
            _lib.call("pfo_cell_forward", ptr(uniq), ptr(n_uniq), u_max, d, c.cell, 1, ptr(GI), None, Gm, ptr(HG),
                      ptr(valid_u), ptr(self.node_feat), ptr(Hnew), ptr(H0))
        else:
            GI = torch.empty(u_max, G, device=dev)
            GH = torch.empty(u_max, G, device=dev)
            _linear(c, ptr(X), ldxin, None, ptr(W_ih), kx, 0, ptr(b_ih), ptr(GI), G, u_max, G, kx, m_dev=ptr(n_uniq))
            _linear(c, ptr(HG), d, None, ptr(W_hh), d, 0, ptr(b_hh), ptr(GH), G, u_max, G, d, m_dev=ptr(n_uniq))
            _lib.call("pfo_cell_forward", ptr(uniq), ptr(n_uniq), u_max, d, c.cell, 0, ptr(GI), ptr(GH), G, ptr(HG),
                      ptr(valid_u), ptr(self.node_feat), ptr(Hnew), ptr(H0))
        tab.update(Hnew=Hnew, lu_u=lu_u, HG=HG, XG=XG, valid_u=valid_u, GI=GI, GH=GH, M1=M1, X2=X2, XH=XH, Wc=Wc,
                   merged=merged, kx=kx, kxp=kxp, ldx=ldx, ldxin=ldxin)
        return tab

    def node_table_backward(self, tab, dH0, g_cell, mlpW=None, g_mlp=None, cellW=None):
        """dH0 (= dHnew) -> gradients of the cell weights and, through the cell input, of the MLP message function
        (memory and stored raw messages are detached inputs)."""
        c, dev = self.cfg, self.device
        d, u_max, n_uniq = c.d, tab["u_max"], tab["n_uniq"]
        G = c.gates * d
        gW_ih, gW_hh, gb_ih, gb_hh = g_cell
        mlp = c.message_fn == "mlp"
        hid, md = c.mlp_hidden, c.msg_dim
        if tab["merged"]:
            kx, kxp, ldx = tab["kx"], tab["kxp"], tab["ldx"]
            Gm = 4 * d if c.cell == CELL_GRU else d
            dG = torch.empty(u_max, Gm, device=dev)
            _lib.call("pfo_cell_backward", ptr(tab["uniq"]), ptr(n_uniq), u_max, d, c.cell, 1, ptr(tab["GI"]), None, Gm,
                      ptr(tab["HG"]), ptr(tab["valid_u"]), ptr(dH0), ptr(dG), None)
            gWc = torch.empty(Gm, ldx, device=dev)
            gbc = torch.empty(Gm, device=dev)
            _wgrad(c, self.ws, ptr(dG), Gm, ptr(tab["XH"]), ldx, None, u_max, Gm, ldx, ptr(gWc), ldx, ptr(gbc),
                   m_dev=ptr(n_uniq))
            _lib.call("pfo_unpack_cell_grads", ptr(gWc), ptr(gbc), d, kx, kxp, c.cell, ptr(gW_ih), ptr(gW_hh),
                      ptr(gb_ih), ptr(gb_hh))
            dGI, W_in, ldw_in, Gin = dG, tab["Wc"], ldx, Gm   # d(cell input) = dG . Wc[:, :kx]
        else:
            dGI = torch.empty(u_max, G, device=dev)
            dGH = torch.empty(u_max, G, device=dev)
            _lib.call("pfo_cell_backward", ptr(tab["uniq"]), ptr(n_uniq), u_max, d, c.cell, 0, ptr(tab["GI"]),
                      ptr(tab["GH"]), G, ptr(tab["HG"]), ptr(tab["valid_u"]), ptr(dH0), ptr(dGI), ptr(dGH))
            X, ldxin, kx = (tab["X2"], tab["ldxin"], md) if mlp else (tab["XG"], c.rawp, c.raw)
            _wgrad(c, self.ws, ptr(dGI), G, ptr(X), ldxin, None, u_max, G, kx, ptr(gW_ih), kx, ptr(gb_ih), m_dev=ptr(n_uniq))
            _wgrad(c, self.ws, ptr(dGH), G, ptr(tab["HG"]), d, None, u_max, G, d, ptr(gW_hh), d, ptr(gb_hh),
                   m_dev=ptr(n_uniq))
            W_in, ldw_in, Gin = (cellW[0] if cellW is not None else None), md, G
        if mlp:
            W1, b1, W2, b2 = mlpW
            gW1, gb1, gW2, gb2 = g_mlp
            M1 = tab["M1"]
            ld1, ld2 = M1.shape[1], (md + 3) // 4 * 4
            dX2 = torch.empty(u_max, ld2, device=dev)
            _linear(c, ptr(dGI), Gin, None, ptr(W_in), ldw_in, 1, None, ptr(dX2), ld2, u_max, md, Gin, m_dev=ptr(n_uniq))
            _wgrad(c, self.ws, ptr(dX2), ld2, ptr(M1), ld1, None, u_max, md, hid, ptr(gW2), hid, ptr(gb2), m_dev=ptr(n_uniq))
            dM1 = torch.empty(u_max, ld1, device=dev)
            _linear(c, ptr(dX2), ld2, None, ptr(W2), hid, 1, None, ptr(dM1), ld1, u_max, hid, md, m_dev=ptr(n_uniq),
                    relu_gate=ptr(M1), ld_gate=ld1)
            _wgrad(c, self.ws, ptr(dM1), ld1, ptr(tab["XG"]), c.rawp, None, u_max, hid, c.raw, ptr(gW1), c.raw, ptr(gb1),
                   m_dev=ptr(n_uniq))

    def _store_overlaps(self):
        """The state update can leave the main stream when it allocates nothing (the `mean` aggregator sorts)."""
        return self.cfg.aggregator == "last"

    def store_state(self, tab, batch, emb, tw, tb, overlap):
        """`persist_and_store`, on the side stream when `overlap` (see `overlap_store`)."""
        if overlap and self.overlap_store and self.side is not None and self._store_overlaps():
            self.side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                self.persist_and_store(tab, batch, emb, tw, tb)
            self._side_pending = True
        else:
            self.persist_and_store(tab, batch, emb, tw, tb)

    def join_store(self):
        if self._side_pending:
            self.join_side()
            self._side_pending = False

    def persist_and_store(self, tab, batch, emb, tw, tb):
        """Persist the positives' memory from the lazy result, then build and store the new raw
        messages with last-wins (tgn.py:185-206)."""
        c, st = self.cfg, self.state
        d, F, B = c.d, c.n_edge_feat, batch["B"]
        src, dst = batch["src"], batch["dst"]
        _lib.call("pfo_persist_rank", ptr(src), ptr(dst), B, d, ptr(st.slot_of_node), ptr(tab["Hnew"]),
                  ptr(st.pend_valid), ptr(st.pend_ts), ptr(st.memory), ptr(st.last_update), ptr(st.last_pos))
        o_src = o_dst = s_src = s_dst = None
        if c.dst_emb_in_msg:    # dyrep: the other endpoint's embedding rides in the message (tgn.py:362-365)
            o_src, o_dst = emb[B:2 * B], emb[:B]
        if c.src_emb_in_msg:    # the node's own embedding in place of its memory row (tgn.py:360-361)
            s_src, s_dst = emb[:B], emb[B:2 * B]
        common = (ptr(st.memory), ptr(st.last_update), ptr(self.edge_feat), ptr(tw), ptr(tb),
                  ptr(o_src), ptr(o_dst), ptr(s_src), ptr(s_dst), ptr(st.pend_msg), c.rawp, ptr(st.pend_ts),
                  ptr(st.pend_valid), ptr(st.last_pos))
        if c.aggregator == "mean":
            # (node, position) pairs of the batch grouped by node, positions ascending inside a node (append order)
            nodes2 = torch.cat([src, dst])
            sorted_node, order = torch.sort(nodes2, stable=True)
            order = order.to(torch.int32)
            _lib.call("pfo_store_messages_mean", ptr(src), ptr(dst), ptr(batch["eidx"]), ptr(batch["ts"]), B, d, F,
                      ptr(sorted_node), ptr(order), *common)
        else:
            _lib.call("pfo_store_messages", ptr(src), ptr(dst), ptr(batch["eidx"]), ptr(batch["ts"]), B, d, F, *common)

    def _slots(self, ids):
        """Node ids -> rows of the batch's node table.  The query list is mapped twice per step (embedding rows and
        the layer-1 query operand): the second request is served from the step's cache."""
        key = (ids.data_ptr(), ids.numel())
        hit = self._slot_cache.get(key)
        if hit is not None and hit[0] is ids:
            return hit[1]
        out = torch.empty(ids.shape, dtype=torch.int32, device=self.device)
        _lib.call("pfo_map_slots", ptr(ids), ids.numel(), self.skip_zero, ptr(self.slot_map), ptr(out))
        self._slot_cache[key] = (ids, out)
        return out

    def _attention_forward(self, tree, W, H0, save):
        """Returns (OUT [M,d], tape)."""
        c = self.cfg
        d, E, Ek, H, ekp, F = c.d, c.E, c.Ek, c.n_heads, c.ekp, c.n_edge_feat
        hd = E // H
        dev = self.device
        layer, M = tree["layer"], tree["M"]
        n = tree["nbr"].shape[1]
        Wqk, cqk, Wc1T, W2, b2 = W[layer - 1]
        tp = _LayerTape()
        tp.layer, tp.M, tp.n = layer, M, n
        tp.child_q = tp.child_n = None
        if layer == 1:
            tp.Tq, tp.qidx = H0, self._slots(tree["nodes"])
            tp.T, tp.idx = H0, self._slots(tree["nbr"].reshape(-1)).view(M, n)
        else:
            out_q, tp.child_q = self._attention_forward(tree["child_q"], W, H0, save)
            out_n, tp.child_n = self._attention_forward(tree["child_n"], W, H0, save)
            ar = torch.arange(M * n, dtype=torch.int32, device=dev).view(M, n)
            tp.Tq, tp.qidx = out_q, torch.arange(M, dtype=torch.int32, device=dev)
            tp.T, tp.idx = out_n, torch.where(tree["nbr"] == 0, torch.full_like(ar, -1), ar)
        tp.eidx, tp.dt = tree["eidx"], tree["dt"]
        ldc, hq = c.ldcat, H * ekp                  # CAT row = [XB (H*ekp) | h_query (d)]
        tp.CAT = torch.empty(M, ldc, device=dev)
        hq_ptr = tp.CAT.data_ptr() + hq * F4
        _lib.call("pfo_gather_rows", ptr(tp.Tq), d, ptr(tp.qidx), M, d, hq_ptr, ldc)
        tp.P = torch.empty(M, H, n, device=dev)
        tp.invalid = torch.empty(M, dtype=torch.int32, device=dev)
        tp.step = layer
        p_drop = c.dropout if save["train"] else 0.0
        U = H0.shape[0]
        if layer == 1 and not save.get("grad", True) and 2 * (U + 1) <= M:
            # qk_h depends on the query NODE only (the query's time encoding is te(0)): when the queries outnumber the
            # rows of the node table -- full-ranking evaluation asks about every stock once per user -- the
            # contraction runs once per table row and the neighbour kernel reads row slot(query node); row U = the
            # operand of a query without a table row (h_q = 0: the bias alone).  Inference only: the backward keeps
            # per-query rows.
            tp.QK = torch.empty(U + 1, H, ekp, device=dev)
            _linear(c, ptr(H0), d, None, ptr(Wqk), d, 0, ptr(cqk), ptr(tp.QK), hq, U, hq, d, m_dev=ptr(save["n_uniq"]))
            tp.QK[U].copy_(cqk.view(H, ekp))
            qk_row = torch.where(tp.qidx < 0, torch.full_like(tp.qidx, U), tp.qidx)
            _lib.call("pfo_attn_nbr_fwd_rows", ptr(tp.QK), ptr(qk_row), ptr(tp.T), d, ptr(tp.idx), ptr(tp.eidx),
                      ptr(tp.dt), ptr(self.edge_feat), ptr(save["tw"]), ptr(save["tb"]), M, n, d, F, H, ekp,
                      float(p_drop), self.seed, tp.step, ptr(self.step_ctr), ptr(tp.CAT), ldc, ptr(tp.P), ptr(tp.invalid))
        else:
            tp.QK = torch.empty(M, H, ekp, device=dev)  # qk_h = Wk_h^T q_h for both heads: one contraction over h_query
            _linear(c, hq_ptr, ldc, None, ptr(Wqk), d, 0, ptr(cqk), ptr(tp.QK), hq, M, hq, d)
            _lib.call("pfo_attn_nbr_fwd", ptr(tp.QK), ptr(tp.T), d, ptr(tp.idx), ptr(tp.eidx), ptr(tp.dt),
                      ptr(self.edge_feat), ptr(save["tw"]), ptr(save["tb"]), M, n, d, F, H, ekp,
                      float(p_drop), self.seed, tp.step, ptr(self.step_ctr), ptr(tp.CAT), ldc, ptr(tp.P), ptr(tp.invalid))
        tp.H1 = torch.empty(M, d, device=dev)       # relu(fc1([out_proj(attn) | h_query])), biases folded into Wc1T
        _linear(c, ptr(tp.CAT), ldc, None, ptr(Wc1T), d, 1, None, ptr(tp.H1), d, M, d, ldc, act=1)
        # the output is NOT kept on the tape: it is the autograd output, and holding it from the
        # backward context would form a reference cycle that only the cyclic GC can free
        out = torch.empty(M, d, device=dev)
        tp.out_rows = M
        _linear(c, ptr(tp.H1), d, None, ptr(W2), d, 0, ptr(b2), ptr(out), d, M, d, d)
        return out, tp

    def _attention_backward(self, tp, dOUT, W, dW, save):
        """Accumulates the folded-operand grads into dW[layer-1] and feature grads into tp.dTq / tp.dT."""
        c = self.cfg
        d, H, ekp, F = c.d, c.n_heads, c.ekp, c.n_edge_feat
        dev = self.device
        M, n = tp.M, tp.n
        ldc, hq = c.ldcat, H * ekp
        Wqk, cqk, Wc1T, W2, b2 = W[tp.layer - 1]
        gWqk, gcqk, gWc1T, gW2, gb2 = dW[tp.layer - 1][:5]
        ws = self.ws
        hq_ptr = tp.CAT.data_ptr() + hq * F4
        # merge MLP: fc2, then fc1 straight down to [dXB | dh_query]
        dH1 = torch.empty(M, d, device=dev)
        _linear(c, ptr(dOUT), d, None, ptr(W2), d, 1, None, ptr(dH1), d, M, d, d, relu_gate=ptr(tp.H1), ld_gate=d)
        _wgrad(c, ws, ptr(dOUT), d, ptr(tp.H1), d, None, M, d, d, ptr(gW2), d, ptr(gb2), accumulate=1)
        dCAT = torch.empty(M, ldc, device=dev)
        _linear(c, ptr(dH1), d, None, ptr(Wc1T), d, 0, None, ptr(dCAT), ldc, M, ldc, d)
        # gWc1T[k, :] = sum_m CAT[m, k] dH1[m, :]  (rows without neighbours carry XB = 0, valid = 0)
        _wgrad(c, ws, ptr(tp.CAT), ldc, ptr(dH1), d, None, M, ldc, d, ptr(gWc1T), d, None, accumulate=1)
        dQK = torch.empty(M, H, ekp, device=dev)
        nws = ws.get(_lib.query("pfo_attn_nbr_bwd_workspace_floats", d))
        p_drop = c.dropout if save["train"] else 0.0
        _lib.call("pfo_attn_nbr_bwd", ptr(tp.QK), ptr(dCAT), ldc, ptr(tp.P), ptr(tp.invalid), ptr(tp.T), d, ptr(tp.idx),
                  ptr(tp.eidx), ptr(tp.dt), ptr(self.edge_feat), ptr(save["tw"]), ptr(save["tb"]),
                  M, n, d, F, H, ekp, float(p_drop), self.seed, tp.step, ptr(self.step_ctr), ptr(dQK), ptr(tp.dT), d,
                  ptr(save["g_twtb"]), 1, ptr(nws))
        dhq_ptr = dCAT.data_ptr() + hq * F4
        _linear(c, ptr(dQK), hq, None, ptr(Wqk), d, 1, None, dhq_ptr, ldc, M, d, hq, accumulate=1)
        _wgrad(c, ws, ptr(dQK), hq, hq_ptr, ldc, None, M, hq, d, ptr(gWqk), d, ptr(gcqk), accumulate=1)
        _lib.call("pfo_scatter_add_rows", dhq_ptr, ldc, ptr(tp.qidx), M, d, ptr(tp.dTq), d)


    # ------------------------------------------------------------------ graph_sum (embedding_module.py:183-219)
    def _sum_forward(self, tree, W, H0, save):
        """GraphSumEmbedding on the neighbour kernels: one head, zero query operand -> uniform softmax weights 1/n
        over the n slots (all live: padded slots resolve to node 0's row), so XB = mean_j x_j and
        relu(sum_j linear_1(x_j)) = relu(n * (W1p . XB + b1)) is the linear kernel's alpha = n.  Returns (OUT, tape)."""
        c = self.cfg
        d, ekp, F = c.d, c.ekp, c.n_edge_feat
        dev = self.device
        layer, M = tree["layer"], tree["M"]
        n = tree["nbr"].shape[1]
        W1p, b1, W2a, b2f = W[layer - 1]
        tp = _LayerTape()
        tp.layer, tp.M, tp.n = layer, M, n
        tp.child_q = tp.child_n = None
        if layer == 1:
            tp.Tq, tp.qidx = H0, self._slots(tree["nodes"])
            tp.T, tp.idx = H0, self._slots(tree["nbr"].reshape(-1)).view(M, n)
        else:
            out_q, tp.child_q = self._sum_forward(tree["child_q"], W, H0, save)
            out_n, tp.child_n = self._sum_forward(tree["child_n"], W, H0, save)
            tp.Tq, tp.qidx = out_q, torch.arange(M, dtype=torch.int32, device=dev)
            tp.T, tp.idx = out_n, torch.arange(M * n, dtype=torch.int32, device=dev).view(M, n)
        tp.eidx, tp.dt = tree["eidx"], tree["dt"]
        tp.QK = torch.zeros(M, 1, ekp, device=dev)
        tp.CAT = torch.empty(M, ekp, device=dev)     # XB = mean of [h | e | te] over the slots (+ psum, valid, one)
        tp.P = torch.empty(M, 1, n, device=dev)
        tp.invalid = torch.empty(M, dtype=torch.int32, device=dev)
        tp.step = layer
        _lib.call("pfo_attn_nbr_fwd", ptr(tp.QK), ptr(tp.T), d, ptr(tp.idx), ptr(tp.eidx), ptr(tp.dt),
                  ptr(self.edge_feat), ptr(save["tw"]), ptr(save["tb"]), M, n, d, F, 1, ekp,
                  0.0, self.seed, tp.step, None, ptr(tp.CAT), ekp, ptr(tp.P), ptr(tp.invalid))
        tp.H1 = torch.empty(M, 2 * d, device=dev)   # [relu(sum_j linear_1(x_j)) | h_query]
        _linear(c, ptr(tp.CAT), ekp, None, ptr(W1p), 2 * d + F, 0, ptr(b1), ptr(tp.H1), 2 * d, M, d, 2 * d + F,
                alpha=float(n), act=1)
        _lib.call("pfo_gather_rows", ptr(tp.Tq), d, ptr(tp.qidx), M, d, tp.H1.data_ptr() + d * F4, 2 * d)
        out = torch.empty(M, d, device=dev)
        tp.out_rows = M
        _linear(c, ptr(tp.H1), 2 * d, None, ptr(W2a), 2 * d, 0, ptr(b2f), ptr(out), d, M, d, 2 * d)
        return out, tp

    def _sum_backward(self, tp, dOUT, W, dW, save):
        c = self.cfg
        d, ekp, F = c.d, c.ekp, c.n_edge_feat
        dev = self.device
        M, n = tp.M, tp.n
        W1p, b1, W2a, b2f = W[tp.layer - 1]
        gW1p, gb1, gW2a, gb2f = dW[tp.layer - 1]
        ws = self.ws
        _wgrad(c, ws, ptr(dOUT), d, ptr(tp.H1), 2 * d, None, M, d, 2 * d, ptr(gW2a), 2 * d, ptr(gb2f), accumulate=1)
        # d/d(pre-activation of the relu) = n * dY gated by Y > 0; d/dh_query = dOUT . W2a[:, d:]
        G1 = torch.empty(M, d, device=dev)
        _linear(c, ptr(dOUT), d, None, ptr(W2a), 2 * d, 1, None, ptr(G1), d, M, d, d, alpha=float(n),
                relu_gate=ptr(tp.H1), ld_gate=2 * d)
        dhq = torch.empty(M, d, device=dev)
        _linear(c, ptr(dOUT), d, None, W2a.data_ptr() + d * F4, 2 * d, 1, None, ptr(dhq), d, M, d, d)
        _wgrad(c, ws, ptr(G1), d, ptr(tp.CAT), ekp, None, M, d, 2 * d + F, ptr(gW1p), 2 * d + F, ptr(gb1), accumulate=1)
        dXB = torch.zeros(M, ekp, device=dev)
        _linear(c, ptr(G1), d, None, ptr(W1p), 2 * d + F, 1, None, ptr(dXB), ekp, M, 2 * d + F, d)
        dQK = torch.empty(M, 1, ekp, device=dev)     # scratch: the query operand is identically zero
        nws = ws.get(_lib.query("pfo_attn_nbr_bwd_workspace_floats", d))
        _lib.call("pfo_attn_nbr_bwd", ptr(tp.QK), ptr(dXB), ekp, ptr(tp.P), ptr(tp.invalid), ptr(tp.T), d, ptr(tp.idx),
                  ptr(tp.eidx), ptr(tp.dt), ptr(self.edge_feat), ptr(save["tw"]), ptr(save["tb"]),
                  M, n, d, F, 1, ekp, 0.0, self.seed, tp.step, None, ptr(dQK), ptr(tp.dT), d,
                  ptr(save["g_twtb"]), 1, ptr(nws))
        _lib.call("pfo_scatter_add_rows", ptr(dhq), d, ptr(tp.qidx), M, d, ptr(tp.dTq), d)


class TGNStepFunction(torch.autograd.Function):
    """One batch of the TGN path as a single autograd node over the packed parameters."""

    @staticmethod
    def forward(ctx, eng: TGNEngine, batch, *flat):
        c, st, dev = eng.cfg, eng.state, eng.device
        d, F = c.d, c.n_edge_feat
        need_grad = any(ctx.needs_input_grad) and batch.get("grad", True)
        it = iter(flat)
        tw, tb = next(it), next(it)
        cellW = [next(it) for _ in range(4)] if c.use_memory else None
        mlpW = [next(it) for _ in range(4)] if c.use_memory and c.message_fn == "mlp" else None
        layerW, embW, rawW, fold_ws = [], None, None, None
        if c.embedding == "graph_attention":
            rawW = [[next(it) for _ in range(10)] for _ in range(c.n_layers)]
            layerW, fold_ws = eng.fold_layers(rawW, tb)             # side stream; joined before the attention
        elif c.embedding == "graph_sum":
            layerW = [[next(it) for _ in range(4)] for _ in range(c.n_layers)]
        elif c.embedding == "time":
            embW = [next(it), next(it)]
        eng.join_store()                                # a forward pass whose backward never ran left its store pending
        q_nodes, q_ts, n, B = batch["q_nodes"], batch["q_ts"], batch["n"], batch["B"]
        Q = q_nodes.shape[0]
        save = dict(train=batch["train"], tw=tw, tb=tb, grad=need_grad)
        eng.step_id += 1
        if c.dropout > 0.0 and batch["train"]:
            eng.step_ctr.add_(16)

        # 1. neighbour sampling tree + unique touched nodes
        tree = None
        id_lists = [q_nodes]
        if c.graph:
            nf = eng.nf
            calls0 = getattr(nf, "call_id", 0)
            with _lib.nvtx_range("K1 neighbour sampling"):
                tree = eng._sample_tree(q_nodes, q_ts, c.n_layers, n, batch.get("q_ids"))
            if getattr(nf, "call_ctr", None) is not None and torch.cuda.is_current_stream_capturing():
                # uniform sampling inside a captured step: the host call ids are baked into the graph, the device
                # counter advances by the step's number of K1 calls on every replay (fresh Philox streams)
                nf.call_ctr.add_(nf.call_id - calls0)
            id_lists = []
            eng._collect_level0(tree, id_lists)
        sb = batch.get("state", batch)                  # interactions that advance the state (default: the embedded ones)
        if sb is not batch and c.use_memory:
            id_lists = id_lists + [sb["src"], sb["dst"]]    # their updated memory rows must be in the node table

        # 2. lazy memory update on the unique nodes (memory_updater.py:35-53, restricted)
        with _lib.nvtx_range("K3 node table (compaction + lazy memory update)"):
            tab = eng.node_table(id_lists, cellW, mlpW) if mlpW is not None else eng.node_table(id_lists, cellW)
        uniq, u_max, n_uniq = tab["uniq"], tab["u_max"], tab["n_uniq"]
        save["n_uniq"] = n_uniq
        H0, Hnew, lu_u = tab["H0"], tab["Hnew"], tab["lu_u"]

        # 3. embeddings
        tape = None
        td = None
        eng._slot_cache = {}                            # slot_of_node was rewritten by this batch's compaction
        qslots = eng._slots(q_nodes)
        eng.qslots_last = qslots                        # rows [src | dst | ...] of the batch in the node table
        eng.q_nodes_last = q_nodes
        if c.embedding == "graph_attention":
            eng.join_side()
            with _lib.nvtx_range("K4 temporal attention forward"):
                emb, tape = eng._attention_forward(tree, layerW, H0, save)
        elif c.embedding == "graph_sum":
            emb, tape = eng._sum_forward(tree, layerW, H0, save)
        elif c.embedding == "time":
            n_src = batch["src"].shape[0]
            emb = torch.empty(Q, d, device=dev)
            td = torch.empty(Q, device=dev)
            ms, ss, md, sd = c.shift
            _lib.call("pfo_time_embedding_fwd", ptr(q_nodes), ptr(q_ts), Q, n_src, d, ptr(eng.slot_map), ptr(Hnew),
                      ptr(lu_u), float(ms), float(ss), float(md), float(sd), ptr(embW[0]), ptr(embW[1]),
                      ptr(td), ptr(emb))
        else:                       # identity: memory'[nodes] (embedding_module.py:32-35)
            emb = torch.empty(Q, d, device=dev)
            _lib.call("pfo_gather_rows", ptr(Hnew), d, ptr(qslots), Q, d, ptr(emb), d)

        # 4. persist positives, then build + store the new raw messages (tgn.py:185-206)
        if c.use_memory and batch["update_state"]:
            with _lib.nvtx_range("K2 persist + message store"):
                eng.store_state(tab, sb, emb, tw, tb, overlap=need_grad and batch["train"])
        out = emb
        if c.use_memory and c.dyrep:    # dyrep returns the updated memory rows (tgn.py:211-215, :322-325)
            out = torch.empty(Q, d, device=dev)
            _lib.call("pfo_gather_rows", ptr(Hnew), d, ptr(qslots), Q, d, ptr(out), d)
        if need_grad:
            ctx.eng, ctx.save = eng, save
            ctx.pack = dict(flat=flat, cellW=cellW, mlpW=mlpW, layerW=layerW, rawW=rawW, fold_ws=fold_ws, embW=embW, tape=tape,
                            tab=tab, u_max=u_max,
                            Hnew=Hnew, qslots=qslots, td=td, Q=Q)
        return out

    @staticmethod
    def backward(ctx, dOut):
        eng, save, pk = ctx.eng, ctx.save, ctx.pack
        c, dev = eng.cfg, eng.device
        d = c.d
        dOut = dOut.contiguous()
        flat = pk["flat"]
        sizes = [t.numel() for t in flat]              # one zero-fill for all operand gradients
        sink = eng.grad_sink
        if sink is not None and sink.numel() == sum(sizes):
            gbuf = sink.zero_()
        else:
            sink, gbuf = None, torch.zeros(sum(sizes), device=dev)
        grads = [g.view_as(t) for g, t in zip(gbuf.split(sizes), flat)]
        it = iter(grads)
        g_tw, g_tb = next(it), next(it)
        g_cell = [next(it) for _ in range(4)] if c.use_memory else None
        g_mlp = [next(it) for _ in range(4)] if c.use_memory and c.message_fn == "mlp" else None
        g_layers, g_raw, g_emb = [], [], None
        if c.embedding == "graph_attention":
            g_raw = [[next(it) for _ in range(10)] for _ in range(c.n_layers)]
            fsz = [t.numel() for t in pk["layerW"][0][:3]]
            fbuf = torch.zeros(c.n_layers * (sum(fsz) + d), device=dev)     # folded-operand grads + the fold's d/dtb
            for l, chunk in enumerate(fbuf.split(sum(fsz) + d)):
                gq, gc, gw, gt = chunk.split(fsz + [d])
                g_layers.append([gq.view_as(pk["layerW"][l][0]), gc, gw.view_as(pk["layerW"][l][2]),
                                 g_raw[l][8], g_raw[l][9], gt])
        elif c.embedding == "graph_sum":
            g_layers = [[next(it) for _ in range(4)] for _ in range(c.n_layers)]
        elif c.embedding == "time":
            g_emb = [next(it), next(it)]
        u_max = pk["u_max"]
        dH0 = torch.zeros(u_max, d, device=dev)          # grad of the unique-node feature table
        attention_grad = c.embedding == "graph_attention" and not (c.use_memory and c.dyrep)
        if c.use_memory and c.dyrep:
            _lib.call("pfo_scatter_add_rows", ptr(dOut), d, ptr(pk["qslots"]), pk["Q"], d, ptr(dH0), d)
        elif c.embedding == "identity":
            _lib.call("pfo_scatter_add_rows", ptr(dOut), d, ptr(pk["qslots"]), pk["Q"], d, ptr(dH0), d)
        elif c.embedding == "time":
            need = 2 * 148 * 2 * d
            buf = eng.ws.get(need)
            gwb = torch.zeros(2 * d, device=dev)
            # slot_of_node may have been rewritten by a later batch: go through the saved slots
            ar = torch.arange(pk["Q"], dtype=torch.int32, device=dev)
            _lib.call("pfo_time_embedding_bwd", ptr(ar), pk["Q"], d, ptr(pk["qslots"]), ptr(pk["Hnew"]), ptr(pk["td"]),
                      ptr(pk["embW"][0]), ptr(pk["embW"][1]), ptr(dOut), ptr(dH0), ptr(gwb), ptr(buf), need)
            g_emb[0].copy_(gwb[:d])
            g_emb[1].copy_(gwb[d:])
        if c.embedding == "graph_sum" and not (c.use_memory and c.dyrep):
            save["g_twtb"] = gbuf[:2 * d]               # [d/dw | d/db] = the first two entries of the flat buffer
            stack = [(pk["tape"], dOut)]
            while stack:
                tp, g = stack.pop()
                if tp.layer == 1:
                    tp.dTq = tp.dT = dH0
                else:
                    tp.dTq = torch.zeros(tp.child_q.out_rows, d, device=dev)
                    tp.dT = torch.zeros(tp.child_n.out_rows, d, device=dev)
                eng._sum_backward(tp, g, pk["layerW"], g_layers, save)
                if tp.layer > 1:
                    stack.append((tp.child_q, tp.dTq))
                    stack.append((tp.child_n, tp.dT))
        if attention_grad:
            save["g_twtb"] = gbuf[:2 * d]               # the neighbour kernels accumulate into the zeroed flat buffer
            # walk the tape from the outermost layer down; level-0 feature grads land in dH0
            stack = [(pk["tape"], dOut)]
            while stack:
                tp, g = stack.pop()
                if tp.layer == 1:
                    tp.dTq = tp.dT = dH0
                else:
                    tp.dTq = torch.zeros(tp.child_q.out_rows, d, device=dev)
                    tp.dT = torch.zeros(tp.child_n.out_rows, d, device=dev)
                with _lib.nvtx_range("K4 temporal attention backward"):
                    eng._attention_backward(tp, g, pk["layerW"], g_layers, save)
                if tp.layer > 1:
                    stack.append((tp.child_q, tp.dTq))
                    stack.append((tp.child_n, tp.dT))
            # the fold's adjoint runs beside the memory-updater backward (it needs only the finished operand grads)
            eng.unfold_layer_grads(pk["rawW"], flat[1], pk["fold_ws"], g_layers, g_raw, [g[5] for g in g_layers])
        with _lib.nvtx_range("K3 memory updater backward"):
            if c.use_memory and g_mlp is not None:
                eng.node_table_backward(pk["tab"], dH0, g_cell, pk["mlpW"], g_mlp, pk["cellW"])
            elif c.use_memory:
                eng.node_table_backward(pk["tab"], dH0, g_cell)
        if attention_grad:
            eng.join_side()
            for g in g_layers:
                g_tb.add_(g[5])
        eng.join_store()                                # the forward pass's state update (side stream) ends here at the latest
        if sink is not None:                            # gradients are in place: nothing for autograd to route
            return (None, None) + (None,) * len(flat)
        return (None, None) + tuple(grads)

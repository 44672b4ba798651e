"""Node-sharded multi-GPU path: one process per GPU, state partitioned by node id.

The reference is single-process (main.py:103); this module ADDS the sharding the north_star asks for:

* owner(x) = x mod G.  Each rank owns, for its nodes only: memory, last_update, the pending-message table and the
  CSR adjacency rows (stored under the local index x // G).  Node and edge features, the price table and the weights
  are replicated (static / small).
* The global batch is split by interaction index (rank r takes the r-th slice); per step the ranks exchange
    R1  (node, t) queries -> owners, K1 on the local CSR rows, neighbour lists back in the same slots;
    R2  unique touched node ids -> owners, lazy GRU/RNN there (de-duplicated over all requesters), updated-memory
        rows + last_update' back in the same slots;
    R3  (training) d(loss)/d(row) back to the owners along R2's slots, cell backward there;
    R4  message rows built where the interactions live -> owners, persist + last-wins there, the winner chosen by
        GLOBAL batch position so the result does not depend on G;
  followed by one all-reduce of the ~133 k parameter gradients (flat bucket, parameter .grad are views into it).
* Every exchange plan lives on the DEVICE (csrc/route_kernels.cu): a row's place in the send buffer is
  slot = dest * cap + arrival order (warp-aggregated atomics), every (source, destination) pair owns `cap` rows, the
  buffers have static shapes, empty slots carry id -1 and are skipped by the consumer kernels, replies come back in the
  same slots.  No split size ever crosses the host, so the whole step -- six exchanges, the all-reduce and Adam included
  -- is captured in ONE CUDA graph per batch size like the single-GPU step.  Capacities start at a size that cannot
  overflow, are calibrated from the bucket counts of the two eager warm-up steps (max over ranks, x margin) and frozen at
  capture; an overflow flag on the device guards them.
* Transport (csrc/peer_kernels.cu): a calibrated exchange is a PUSH of each destination's block straight into that
  rank's receive buffer -- a cudaMalloc'ed arena every rank maps through CUDA IPC, same offset everywhere -- followed by
  a flag barrier over the same peer mappings (release / acquire at system scope, epochs in device memory, timeout
  instead of a hang).  Two small launches instead of one NCCL all-to-all (1.185 -> 1.111 ms per step on 2 GPUs,
  1.244 -> 1.180 ms on 8).  The oversized calibration steps, and a box without peer access, use all_to_all_single.
* R4 runs on the engine's side stream beside the BPR loss and the attention backward (nothing it writes is read there).
* Every quantity is identical to the 1-GPU path on the same global batch up to fp32 summation order
  (tools/check_sharded.py under torchrun, tests/test_gpu_sharded.py: loss, gradients, memory, last_update, pending
  flags, evaluation candidates and scores, on 2 GPUs; the same machinery on one rank runs in the 1-GPU test tier).
* A procedural `synth_device.DeviceStream` (BASELINE config 4: 10^9 interactions) is consumed without ever existing on
  the host: columns per batch from the interaction index, CSR rows of the owned nodes from 64 M-interaction chunks.

`Exchange` also runs on CPU tensors with the gloo backend (plans by torch ops instead of the CUDA kernels), which is
how the slot / reply logic is covered by world_size-2 tests without a GPU (tests/test_dist_router.py).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr
from .engine import TGNEngine, TGNState, ModelConfig, _linear
from .graph import TemporalCSR, NeighborFinder
from .trainer import PfoTrainer, allreduce_sum_, replica_slice
from .synth_device import DeviceStream

F4 = 4


class Plan:
    __slots__ = ("slot", "local", "cap", "rows", "counts", "calibrated")


class _RawCuda:
    """A raw device allocation as an object torch.as_tensor can wrap without copying (__cuda_array_interface__)."""

    def __init__(self, address, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(address), False),
                                         "version": 2}


class PeerArena:
    """One cudaMalloc'ed arena per rank, mapped into every other rank of the node through CUDA IPC
    (csrc/peer_kernels.cu).  Receive buffers are bump-allocated from it: the ranks run the same sequence of exchanges
    with the same (calibrated) shapes, so a buffer has the same offset in every arena and a peer can store straight
    into it.  `begin_step` rewinds the allocator and places the step-start barrier."""

    def __init__(self, device, group, nbytes):
        import ctypes
        self.device, self.group = torch.device(device), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > int(_lib.query("pfo_peer_max_ranks")):
            raise _lib.PfoError("peer transport: too many ranks for one node's arena table")
        self.nbytes = int(nbytes)
        lib = _lib.load()
        ptr_, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        rc = lib.pfo_peer_alloc(self.nbytes, ctypes.byref(ptr_), handle)
        if rc != 0:
            raise _lib.PfoError(f"pfo_peer_alloc failed with cudaError {rc}")
        self.local = int(ptr_.value)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        bases = []
        for g in range(self.world):
            if g == self.rank:
                bases.append(self.local)
                continue
            p, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64).from_buffer_copy(handles[g])
            rc = lib.pfo_peer_open(h, ctypes.byref(p))
            if rc != 0:
                raise _lib.PfoError(f"pfo_peer_open(rank {g}) failed with cudaError {rc}")
            bases.append(int(p.value))
        self.bases = (ctypes.c_uint64 * self.world)(*bases)
        self.bytes_view = torch.as_tensor(_RawCuda(self.local, self.nbytes), device=self.device)
        self.header = (int(_lib.query("pfo_peer_header_bytes")) + 255) // 256 * 256
        self.max_barriers = int(_lib.query("pfo_peer_max_barriers"))
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.offset, self.next_id = self.header, 1
        self.timeout_s = 20.0

    def begin_step(self):
        """Rewind the allocator; barrier 0 keeps a fast rank out of buffers a slow rank still reads from the last step."""
        self.offset, self.next_id = self.header, 1
        _lib.call("pfo_peer_barrier", self.bases, self.world, self.rank, 0, ptr(self.error), self.timeout_s)

    def room(self, nbytes):
        return self.offset + (int(nbytes) + 255) // 256 * 256 <= self.nbytes and self.next_id < self.max_barriers

    def alloc(self, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        off = self.offset
        self.offset += (nbytes + 255) // 256 * 256
        return off, self.bytes_view[off:off + nbytes].view(dtype).view(*shape)

    def all_to_all(self, buf):
        """Equal-split all-to-all of `buf` ([G * rows, w] of 32-bit elements): push + barrier."""
        G = self.world
        off, out = self.alloc(tuple(buf.shape), buf.dtype)
        block_words = buf.numel() // G
        _lib.call("pfo_peer_push", ptr(buf), self.bases, off, G, self.rank, block_words)
        _lib.call("pfo_peer_barrier", self.bases, G, self.rank, self.next_id, ptr(self.error), self.timeout_s)
        self.next_id += 1
        return out


class Exchange:
    """Fixed-capacity bucket exchange with device-side plans (see the module docstring)."""

    def __init__(self, device, group=None, margin=1.4, quantum=256, transport="peer", arena_bytes=2 << 30):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.margin, self.quantum = float(margin), int(quantum)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.batch = (1, 1)         # (this rank's interactions, the largest slice of any rank) of the current batch
        self.frozen = {}            # key -> capacity fixed for the captured graphs
        self.observed = {}          # key -> largest bucket count seen in eager steps (max over ranks)
        self._pending = []          # (key, counts tensor) of the current eager step
        # transport of the all-to-alls: "peer" = direct stores into the other GPUs' arenas + flag barrier for every
        # exchange whose capacity is calibrated (the captured steps); NCCL all_to_all_single otherwise / as fallback
        self.arena = None
        self.transport = "nccl"
        if transport == "peer" and self.device.type == "cuda":
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            try:
                self.arena = PeerArena(self.device, group, arena_bytes)
            except Exception as exc:                         # no IPC / no peer access on this box: NCCL carries everything
                ok.zero_()
                self._peer_error = str(exc)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.transport = "peer"
            else:
                self.arena = None

    # ---- capacities
    def safe_cap(self, rows):
        """A capacity that cannot overflow by construction at G <= 2 and is generous beyond: 3 * rows / G (the eager
        calibration steps run with it; hot stocks skew the buckets by well under 2x at Zipf(0.8))."""
        G = self.world
        return int(rows) if G <= 2 else min(int(rows), int(-(-3 * int(rows) // G)))

    def cap_for(self, key, rows):
        cap = self.frozen.get(key)
        return cap if cap is not None else max(self.safe_cap(rows), 1)

    def common_rows(self, rows):
        """The row count the RANK WITH THE LARGEST SLICE has for this exchange.  Buffer shapes must agree on every rank
        (equal-split all-to-all), but the short last batch of an epoch gives the ranks slices that differ by one
        interaction: every exchange carries a whole number of rows per interaction, so it is scaled to the largest
        slice -- a pure function of the global batch, identical everywhere."""
        b_loc, b_max = self.batch
        return int(rows) if b_loc == b_max else -(-int(rows) // b_loc) * b_max

    def collect(self):
        """End of an eager step: one host read of the bucket counts, max over ranks (a collective: every rank calls it
        after every eager step).  Never called during graph capture or replay."""
        if not self._pending:
            return
        keys = [k for k, _ in self._pending]
        mx = torch.stack([c.max() for _, c in self._pending]).to(torch.int64)
        if self.world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        for k, v in zip(keys, mx.tolist()):
            self.observed[k] = max(self.observed.get(k, 0), int(v))
        self._pending = []

    def freeze(self):
        """Fix the capacity of every exchange observed so far (called right before a graph capture)."""
        for k, v in self.observed.items():
            if k not in self.frozen:
                q = self.quantum
                self.frozen[k] = max(q, int(-(-int(v * self.margin + 1) // q) * q))

    def begin_step(self):
        """Start of a (train or evaluation) step, inside the captured region."""
        if self.arena is not None:
            self.arena.begin_step()

    def check_overflow(self):
        """Host read of the overflow flag (one sync): raised when a frozen capacity turned out too small."""
        flag = self.overflow.clone()
        if self.arena is not None:
            flag |= self.arena.error
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        v = int(flag.item())
        if v & 2:
            raise RuntimeError("peer-memory barrier timed out: a rank did not reach an exchange (results are invalid)")
        if v & 1:
            raise RuntimeError("node-sharded exchange overflowed a calibrated bucket capacity: rows were dropped; "
                               "re-run with a larger `exchange_margin`")

    # ---- plan + movement
    def plan(self, tag, ids, rows, n_valid=None, cap_rows=None):
        """ids int32[rows] (global node ids; < 0 = no row).  Returns the plan: slot / owner-local id per row.  The
        capacity (and the calibration key) depend on `cap_rows`, the rank-independent size of this exchange."""
        G = self.world
        p = Plan()
        cap_rows = self.common_rows(rows) if cap_rows is None else int(cap_rows)
        key = tuple(tag) + (cap_rows,)
        p.rows, p.cap = int(rows), self.cap_for(key, cap_rows)
        p.calibrated = key in self.frozen
        dev = ids.device
        p.counts = torch.empty(G, dtype=torch.int32, device=dev)
        p.slot = torch.empty(rows, dtype=torch.int32, device=dev)
        p.local = torch.empty(rows, dtype=torch.int32, device=dev)
        if dev.type == "cuda":
            _lib.call("pfo_route_plan", ptr(ids), rows, ptr(n_valid), G, p.cap, ptr(p.counts), ptr(p.slot), ptr(p.local),
                      ptr(self.overflow))
        else:                       # host logic of the gloo tests: the same plan by torch ops (stable arrival order)
            idl = ids.long()
            ok = idl >= 0
            if n_valid is not None:
                ok &= torch.arange(rows) < int(n_valid.item())
            dest = torch.where(ok, idl % G, torch.full_like(idl, G))
            onehot = torch.nn.functional.one_hot(dest, G + 1)[:, :G]
            pos = (torch.cumsum(onehot, 0) - onehot).gather(1, dest.clamp(max=G - 1).view(-1, 1)).view(-1)
            p.counts.copy_(onehot.sum(0).to(torch.int32))
            fits = ok & (pos < p.cap)
            if bool((ok & ~fits).any()):
                self.overflow.fill_(1)
            p.slot.copy_(torch.where(fits, dest * p.cap + pos, torch.full_like(pos, -1)).to(torch.int32))
            p.local.copy_(torch.where(ok, idl // G, torch.full_like(idl, -1)).to(torch.int32))
        if key not in self.frozen and not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            self._pending.append((key, p.counts))
        return p

    def buffer(self, plan, width, dtype=torch.int32, fill=None):
        buf = torch.empty(self.world * plan.cap, width, dtype=dtype, device=plan.slot.device)
        if fill is not None:
            buf.fill_(fill)
        return buf

    def scatter(self, plan, rows, buf, n_valid=None):
        """buf[slot[i], :w] = rows[i, :w] (32-bit words of any type; rows may be a column block of a wider tensor);
        n_valid (device int32, optional): only the first *n_valid rows are walked."""
        w = rows.shape[1] if rows.dim() > 1 else 1
        if rows.device.type == "cuda":
            _lib.call("pfo_scatter_rows", ptr(rows), rows.stride(0) if rows.dim() > 1 else 1, ptr(plan.slot), plan.rows,
                      ptr(n_valid), w, ptr(buf), buf.stride(0))
        else:
            ok = plan.slot >= 0
            buf.view(buf.shape[0], -1)[plan.slot[ok].long(), :w] = rows.view(plan.rows, -1)[ok]
        return buf

    def gather(self, plan, buf, col0, w, out, fill=0):
        """out[i, :w] = buf[slot[i], col0:col0+w], `fill` (a 32-bit pattern) where the row was dropped."""
        if buf.device.type == "cuda":
            _lib.call("pfo_gather_words", buf.data_ptr() + col0 * F4, buf.stride(0), ptr(plan.slot), plan.rows, w,
                      ptr(out), out.stride(0) if out.dim() > 1 else 1, int(fill) & 0xffffffff)
        else:
            ok = plan.slot >= 0
            o = out.view(plan.rows, -1)
            o.fill_(0)
            o[ok] = buf[plan.slot[ok].long(), col0:col0 + w]
        return out

    def all_to_all(self, buf, plan=None):
        """Equal-split all-to-all of a slot buffer.  Calibrated exchanges (fixed, tight capacities: every captured step)
        go through the peer arena when it has room; the oversized buffers of the calibration steps go through NCCL."""
        if (self.arena is not None and plan is not None and plan.calibrated and buf.is_contiguous()
                and buf.element_size() == 4 and self.arena.room(buf.numel() * 4)):
            return self.arena.all_to_all(buf)
        out = torch.empty_like(buf)
        if self.world > 1:
            dist.all_to_all_single(out, buf, group=self.group)
        else:
            out.copy_(buf)
        return out


def local_csr(st_sources, st_destinations, st_edge_idxs, st_timestamps, n_nodes, rank, world, device):
    """CSR rows of the nodes owned by `rank`, indexed by x // world; neighbour ids stay global.  Built on the device
    when `device` is CUDA: the 2E (node, neighbour, edge, time) entries are filtered to the owned nodes first, then
    ordered by one stable sort by time and one stable sort by local node (= the per-node stable sort of
    utils/utils.py:139), so a rank only ever sorts its own share."""
    dev = torch.device(device)
    n_local = (n_nodes + world - 1) // world
    csr = TemporalCSR.__new__(TemporalCSR)
    csr.n_nodes, csr.n_events, csr.device = n_local, int(len(st_sources)), dev
    if dev.type == "cuda":
        src = torch.as_tensor(np.asarray(st_sources, dtype=np.int64), device=dev)
        dst = torch.as_tensor(np.asarray(st_destinations, dtype=np.int64), device=dev)
        eid = torch.as_tensor(np.asarray(st_edge_idxs, dtype=np.int64), device=dev)
        ts = torch.as_tensor(np.asarray(st_timestamps, dtype=np.float64), device=dev)
        return _local_csr_device(csr, src, dst, eid, ts, n_local, rank, world)
    src = np.asarray(st_sources, dtype=np.int64)
    dst = np.asarray(st_destinations, dtype=np.int64)
    node = np.stack([src, dst], axis=1).ravel()
    other = np.stack([dst, src], axis=1).ravel()
    eid = np.repeat(np.asarray(st_edge_idxs, dtype=np.int64), 2)
    ts = np.repeat(np.asarray(st_timestamps, dtype=np.float64), 2)
    mine = np.nonzero(node % world == rank)[0]
    ln = node[mine] // world
    order = np.lexsort((np.arange(mine.shape[0]), ts[mine], ln))
    rowptr = np.zeros(n_local + 1, dtype=np.int64)
    np.cumsum(np.bincount(ln, minlength=n_local), out=rowptr[1:])
    csr.rowptr = torch.as_tensor(rowptr, device=dev)
    csr.nbr = torch.as_tensor(other[mine][order].astype(np.int32), device=dev)
    csr.eidx = torch.as_tensor(eid[mine][order].astype(np.int32), device=dev)
    csr.ts = torch.as_tensor(ts[mine][order], device=dev)
    return csr


def _local_csr_device(csr, src, dst, eid, ts, n_local, rank, world):
    """Device build of `local_csr` from device columns."""
    return build_local_csr(csr, [(0, src, dst)], lambda ev: ts[ev], lambda ev: eid[ev], n_local, rank, world)


def build_local_csr(csr, chunks, ts_of, eidx_of, n_local, rank, world):
    """CSR rows of the owned nodes from an iterator of (first interaction index, sources, destinations) device chunks --
    a 10^9-interaction stream never sits in memory at once: each chunk is filtered to the entries whose node this rank
    owns (each interaction is appended to both endpoints, utils/utils.py:122-123), only those are kept, and timestamps
    / edge idxs are looked up for the survivors (`ts_of`, `eidx_of`: functions of the interaction index).  Order inside
    a row = (timestamp, stream order) through two stable device sorts (torch.sort on CUDA is a radix sort)."""
    ev_l, ln_l, other_l, side_l = [], [], [], []
    dev = None
    for i0, src, dst in chunks:
        dev = src.device
        src, dst = src.long(), dst.long()
        for side, (node, other) in enumerate(((src, dst), (dst, src))):
            mine = torch.nonzero(node % world == rank).view(-1)
            ev_l.append(mine + i0)
            ln_l.append(torch.div(node[mine], world, rounding_mode="floor").to(torch.int32))
            other_l.append(other[mine].to(torch.int32))
            side_l.append(torch.full_like(mine, side, dtype=torch.int8))
    ev, ln, other, side = torch.cat(ev_l), torch.cat(ln_l), torch.cat(other_l), torch.cat(side_l)
    del ev_l, ln_l, other_l, side_l
    # stream order of the 2E entries is (interaction, side): restore it, then sort by time, then by node (both stable)
    o0 = torch.sort(ev * 2 + side.long()).indices
    ev, ln, other = ev[o0], ln[o0], other[o0]
    del o0, side
    t = ts_of(ev)
    o1 = torch.sort(t, stable=True).indices
    o2 = torch.sort(ln[o1], stable=True).indices
    order = o1[o2]
    del o1, o2
    csr.nbr = other[order].contiguous()
    csr.eidx = eidx_of(ev[order]).to(torch.int32).contiguous()
    csr.ts = t[order].contiguous()
    counts = torch.bincount(ln.long(), minlength=n_local)
    csr.rowptr = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=csr.rowptr[1:])
    return csr


def local_csr_from_device_stream(ds, n_events, rank, world, chunk=1 << 26):
    """`local_csr` of the first `n_events` interactions of a procedural `synth_device.DeviceStream`."""
    n_local = (ds.n_nodes + world - 1) // world
    csr = TemporalCSR.__new__(TemporalCSR)
    csr.n_nodes, csr.n_events, csr.device = n_local, int(n_events), ds.device

    def chunks():
        for i0 in range(0, n_events, chunk):
            i = torch.arange(i0, min(n_events, i0 + chunk), dtype=torch.int64, device=ds.device)
            src, dst = ds.src_dst(i)
            yield i0, src, dst

    return build_local_csr(csr, chunks(), ds.timestamps, lambda ev: ev + 1, n_local, rank, world)


class ShardedNeighborFinder(NeighborFinder):
    """K1 over the owners' CSR rows: exchange R1.  `sample` has the interface of `NeighborFinder.sample`; the queries
    travel to the rank that owns their row as [local node id | t | query id], K1 runs there over every slot of the
    receive buffer (empty slots answer with zero rows) writing [nbr | eidx | dt] straight into the reply row, and the
    requester picks its answers out of the same slots."""

    def __init__(self, local_csr_, exchange: Exchange, uniform=False, seed=0, tag="nf"):
        super().__init__(local_csr_, uniform=uniform, seed=seed)
        self.ex, self.tag = exchange, tag

    def sample(self, q_nodes, q_ts, n_neighbors, out=None, q_ids=None, ld_out=0):
        ex, Q = self.ex, q_nodes.shape[0]
        n = max(int(n_neighbors), 1)
        dev = q_nodes.device
        if n_neighbors <= 0:
            return super().sample(q_nodes, q_ts, n_neighbors)
        plan = ex.plan((self.tag, "R1"), q_nodes, Q)
        req = ex.buffer(plan, 4, fill=-1)
        _lib.call("pfo_pack_queries", ptr(plan.local), ptr(q_ts), ptr(q_ids), ptr(plan.slot), Q, ptr(req))
        got = ex.all_to_all(req, plan)
        R = got.shape[0]
        qn = torch.empty(R, dtype=torch.int32, device=dev)
        qt = torch.empty(R, dtype=torch.float64, device=dev)
        qi = torch.empty(R, dtype=torch.int32, device=dev)
        _lib.call("pfo_unpack_queries", ptr(got), R, ptr(qn), ptr(qt), ptr(qi))
        reply = torch.empty(R, 3 * n, dtype=torch.int32, device=dev)
        rf = reply.view(torch.float32)
        NeighborFinder.sample(self, qn, qt, n, out=(reply[:, :n], reply[:, n:2 * n], None, rf[:, 2 * n:]),
                              q_ids=qi if self.uniform else None, ld_out=3 * n)
        back = ex.all_to_all(reply, plan)
        nbr = torch.empty(Q, n, dtype=torch.int32, device=dev)
        eidx = torch.empty(Q, n, dtype=torch.int32, device=dev)
        dt = torch.empty(Q, n, dtype=torch.float32, device=dev)
        _lib.call("pfo_unroute_neighbors", ptr(back), ptr(plan.slot), Q, n, ptr(nbr), ptr(eidx), ptr(dt))
        return nbr, eidx, None, dt


class _Scratch:
    """Compaction scratch over the GLOBAL id space (requester side)."""

    def __init__(self, n_nodes, device):
        self.bitmap = torch.zeros((n_nodes + 31) // 32, dtype=torch.int32, device=device)
        self.slot_of_node = torch.zeros(n_nodes, dtype=torch.int32, device=device)
        self.compact_ws = torch.zeros(int(_lib.query("pfo_compact_workspace_ints", n_nodes)), dtype=torch.int32,
                                      device=device)
        self.n_unique = torch.zeros(1, dtype=torch.int32, device=device)


class ShardedEngine(TGNEngine):
    """TGNEngine whose node table, its backward and the message store go through the owners."""

    def __init__(self, cfg: ModelConfig, state: TGNState, node_feat, edge_feat, nf, exchange: Exchange,
                 n_nodes_global):
        if cfg.message_fn != "identity" or cfg.aggregator != "last" or cfg.src_emb_in_msg:
            raise NotImplementedError("the node-sharded engine covers the models main.py builds: identity message "
                                      "function, `last` aggregator, memory rows in the messages")
        self.ex, self.G, self.rank = exchange, exchange.world, exchange.rank
        self.n_global = int(n_nodes_global)
        super().__init__(cfg, state, node_feat, edge_feat, nf)
        self.req = _Scratch(self.n_global, node_feat.device)
        self.tag = "train"
        # R4 beside the loss and the attention backward (side stream): parity-green on 2 GPUs, but measured SLOWER there
        # (1.066 against 1.038 ms/step, profiles/r2_bench_2gpu_sharded_r4_overlap.json: the exchange and the backward
        # kernels compete for the same SMs and the join sits on the critical path), so it is an opt-in
        self.overlap_store = os.environ.get("PFO_SHARDED_OVERLAP", "0") == "1"
        self._side_pending = False

    def _query_ids(self, groups, B):
        """Position of each query in the query list of the un-sharded step on the global batch: group g holds
        k_g * B_global rows interaction-major and this rank's slice starts at interaction key_base of the batch."""
        if not getattr(self.nf, "uniform", False):
            return None
        dev, out, off = self.device, [], 0
        base, Bg = getattr(self, "key_base", self.rank * B), getattr(self, "key_side", B * self.G)
        for g in groups:
            k = g.shape[0] // B
            out.append(off + base * k + torch.arange(g.shape[0], dtype=torch.int32, device=dev))
            off += Bg * k
        return torch.cat(out)

    def node_table(self, id_lists, cellW, mlpW=None):
        c, st, dev, G, ex = self.cfg, self.state, self.device, self.G, self.ex
        d = c.d
        uniq, u_max = self._unique_nodes(id_lists, scratch=self.req, n_nodes=self.n_global)
        n_uniq = self.req.n_unique.clone()
        self.slot_map = self.req.slot_of_node
        if not c.use_memory:
            return dict(uniq=uniq, u_max=u_max, n_uniq=n_uniq, H0=self.node_feat.index_select(0, uniq.long()),
                        Hnew=None, lu_u=None)
        # R2: unique ids -> owners
        total = sum(int(ids.numel()) for ids in id_lists)
        plan = ex.plan((self.tag, "R2"), uniq, u_max, n_valid=n_uniq, cap_rows=min(ex.common_rows(total), self.n_global))
        req = ex.buffer(plan, 1, fill=-1)
        ex.scatter(plan, plan.local.view(-1, 1), req, n_valid=n_uniq)
        got = ex.all_to_all(req, plan).view(-1)
        R = got.shape[0]
        _lib.call("pfo_mark_nodes", ptr(got), R, 0, ptr(st.bitmap))
        u_own = min(max(R, 1), st.n_nodes)
        uo = torch.zeros(u_own, dtype=torch.int32, device=dev)
        _lib.call("pfo_compact_nodes", ptr(st.bitmap), st.n_nodes, ptr(st.compact_ws), ptr(uo),
                  ptr(st.slot_of_node), ptr(st.n_unique))
        n_own = st.n_unique.clone()
        # the memory updater on the owned rows (H0 = memory' + node features is formed on the requester: the owner's
        # ids are local indices, its H0 output is scratch)
        own = self._cell_table(uo, n_own, u_own, cellW)
        Hnew_own, lu_own = own["Hnew"], own["lu_u"]
        slots_own = torch.empty(R, dtype=torch.int32, device=dev)
        _lib.call("pfo_map_slots", ptr(got), R, 0, ptr(st.slot_of_node), ptr(slots_own))
        reply = torch.empty(R, d + 1, device=dev)           # [updated memory row | last_update'] per received id
        _lib.call("pfo_route_reply_rows", ptr(Hnew_own), ptr(lu_own), ptr(slots_own), R, d, ptr(reply))
        back = ex.all_to_all(reply, plan)
        Hnew = torch.empty(u_max, d, device=dev)
        lu_u = torch.empty(u_max, device=dev)
        H0 = torch.empty(u_max, d, device=dev)              # rows of the table + node features, in one pass
        _lib.call("pfo_unroute_rows", ptr(back), ptr(plan.slot), ptr(uniq), ptr(n_uniq), ptr(self.node_feat), u_max, d,
                  ptr(Hnew), ptr(lu_u), ptr(H0))
        own.update(slots=slots_own, R=R)
        return dict(uniq=uniq, u_max=u_max, n_uniq=n_uniq, H0=H0, Hnew=Hnew, lu_u=lu_u, plan=plan, own=own, n_req=n_uniq)

    def node_table_backward(self, tab, dH0, g_cell, mlpW=None, g_mlp=None, cellW=None):
        # R3: gradient rows -> owners along R2's slots, summed over requesters, then the cell backward of the base class
        own, ex, d = tab["own"], self.ex, self.cfg.d
        self.join_store()               # the message exchange of the forward pass: its collective was issued first
        send = ex.buffer(tab["plan"], d, dtype=torch.float32)
        ex.scatter(tab["plan"], dH0, send, n_valid=tab["n_req"])
        got = ex.all_to_all(send, tab["plan"])
        dH_own = torch.zeros(own["u_max"], d, device=self.device)
        _lib.call("pfo_scatter_add_rows", ptr(got), d, ptr(own["slots"]), own["R"], d, ptr(dH_own), d)
        TGNEngine.node_table_backward(self, own, dH_own, g_cell)

    def persist_and_store(self, tab, batch, emb, tw, tb):
        """R4: message rows built here, applied at the owners with last-wins by global batch position.  In a training
        step nothing downstream of the forward pass reads the state this writes, so the whole exchange runs on the
        engine's side stream, beside the BPR loss and the backward pass, and is joined at the end of the backward."""
        B = batch["B"]
        qs = self.qslots_last                               # the query list starts with [src | dst]: their table rows
        self._persist_and_store(tab, batch, emb, tw, tb, qs[:B], qs[B:2 * B])   # stream chosen by TGNEngine.store_state

    def _store_overlaps(self):
        return True                                         # R4's buffers belong to the stream that allocates them

    def _persist_and_store(self, tab, batch, emb, tw, tb, s_slot, d_slot):
        c, st, dev, G, ex = self.cfg, self.state, self.device, self.G, self.ex
        d, F, B = c.d, c.n_edge_feat, batch["B"]
        src, dst = batch["src"], batch["dst"]
        ldr = (c.raw + 3 + 3) // 4 * 4                       # message | owner-local node | global position | fp32 time
        rows = torch.empty(2 * B, ldr, device=dev)
        o_src = o_dst = None
        if c.dst_emb_in_msg:
            o_src, o_dst = emb[B:2 * B].contiguous(), emb[:B].contiguous()
        key_base = int(getattr(self, "key_base", self.rank * B))
        key_side = int(getattr(self, "key_side", B * G))
        _lib.call("pfo_build_routed_messages", ptr(s_slot), ptr(d_slot), ptr(src), ptr(dst), ptr(batch["eidx"]),
                  ptr(batch["ts"]), B, d, F, ptr(tab["Hnew"]), ptr(tab["lu_u"]), ptr(self.edge_feat), ptr(tw), ptr(tb),
                  ptr(o_src), ptr(o_dst), G, key_base, key_side, ptr(rows), ldr)
        nodes = self.q_nodes_last[:2 * B]                    # the query list starts with [src | dst]
        plan = ex.plan((self.tag, "R4"), nodes, 2 * B)
        send = ex.buffer(plan, ldr, dtype=torch.float32)
        send.view(torch.int32)[:, c.raw].fill_(-1)           # empty slots: node id -1 (only this column is read first)
        ex.scatter(plan, rows, send)
        got = ex.all_to_all(send, plan)
        own = tab["own"]
        _lib.call("pfo_apply_routed_messages", ptr(got), ldr, got.shape[0], d, c.raw, ptr(st.slot_of_node),
                  ptr(own["Hnew"]), ptr(st.memory), ptr(st.last_update), ptr(st.pend_msg), c.rawp, ptr(st.pend_ts),
                  ptr(st.pend_valid), ptr(st.last_pos))


class ShardedTrainer(PfoTrainer):
    """PfoTrainer over G ranks with node-sharded state: global batch [s, e), rank r trains / evaluates its r-th slice."""

    def __init__(self, st, tc, device, rank, world, group=None, exchange_margin=1.4, nccl_in_graph=True, transport=None):
        import os
        self.rank, self.world, self.group = int(rank), int(world), group
        _lib.use_device(device)
        self.ex = Exchange(device, group=group, margin=exchange_margin,
                           transport=transport or os.environ.get("PFO_TRANSPORT", "peer"))
        if tc.model == "dyrep":
            raise NotImplementedError("dyrep messages carry embeddings: routed, but not verified in this mode")
        super().__init__(st, tc, device)
        self._loss_scale = 1.0 / self.world
        self.nccl_in_graph = bool(nccl_in_graph)
        self._graph_tail_eager = not self.nccl_in_graph
        # NCCL's watchdog thread touches the CUDA API while this thread captures: keep the capture check thread-local
        self._capture_kw = {"capture_error_mode": "thread_local"}
        self.engine.seed = tc.seed + 7919 * self.rank         # decorrelate the dropout streams of the ranks
        self.params = [p for p in self.tgn.parameters() if p.requires_grad]
        self._own_bucket = self.gflat is None                 # else: the flat gradient buffer of the base class (engine sink)
        if self._own_bucket:
            sizes = [p.numel() for p in self.params]
            self.gflat = torch.zeros(sum(sizes), device=self.device)
            for p, g in zip(self.params, self.gflat.split(sizes)):
                p.grad = g.view_as(p)

    # ---- construction hooks
    @property
    def procedural(self):
        return isinstance(self.st, DeviceStream)

    def _prepare_stream(self, train_frac_mask=None):
        """Host streams go through the base class.  A procedural `DeviceStream` (scale configuration, 10^9 interactions)
        never materialises on the host: the split point comes from a bisection on its monotone timestamps, the feature
        tables are drawn on the device, and batches are evaluated from the interaction index on demand."""
        if not self.procedural:
            return super()._prepare_stream(train_frac_mask)
        st, tc, dev = self.st, self.tc, self.device
        if tc.model == "jodie":
            raise NotImplementedError("the time embedding needs per-node inter-event statistics of the whole stream "
                                      "(utils/data.py:75-99): not computed for procedural streams")
        self.n_train, self.masks = st.n_train(0.8), None
        self._n_val_end = st.n_train(0.9)
        self.dev_stream = None
        self._col_cache = {}
        g = torch.Generator(device=dev).manual_seed(0)
        node_feat = torch.rand(st.n_nodes, tc.d, device=dev, generator=g)          # main.py:87: U(0, 1)
        edge_feat = torch.zeros(st.n_events + 1, 1, device=dev)
        present = torch.zeros(st.n_items, dtype=torch.bool, device=dev)
        present_tr = torch.zeros(st.n_items, dtype=torch.bool, device=dev)
        chunk = 1 << 26
        for i0 in range(0, st.n_events, chunk):
            i = torch.arange(i0, min(st.n_events, i0 + chunk), dtype=torch.int64, device=dev)
            edge_feat[i0 + 1:i0 + 1 + i.numel(), 0] = st.edge_feature(i)
            item = st.src_dst(i)[1].long() - st.n_users - 1
            present[item] = True
            n_in = max(0, min(self.n_train - i0, i.numel()))
            if n_in > 0:
                present_tr[item[:n_in]] = True
        off = st.n_users + 1
        return dict(train_index=None, node_feat=node_feat, edge_feat=edge_feat, time_statistics=(0.0, 1.0, 0.0, 1.0),
                    universe_train=torch.nonzero(present_tr).view(-1).cpu().numpy() + off,
                    universe_all=torch.nonzero(present).view(-1).cpu().numpy() + off)

    def _columns(self, s, e):
        """Device columns of interactions [s, e) of a procedural stream (small cache: a batch is usually asked for twice,
        by the static-buffer fill and by the host-batch builder)."""
        key = (int(s), int(e))
        c = self._col_cache.get(key)
        if c is None:
            if len(self._col_cache) >= 256:
                self._col_cache.pop(next(iter(self._col_cache)))
            c = self._col_cache[key] = self.st.columns(s, e)
        return c

    def prefetch(self, spans):
        """Evaluate the columns of the given (s, e) GLOBAL batches ahead of time (this rank's slices), so that a timed
        loop only copies them into the static buffers."""
        if self.procedural:
            for s, e in spans:
                self._columns(*replica_slice(s, e, self.rank, self.world))

    def _batch(self, s, e):
        if not self.procedural:
            return super()._batch(s, e)
        c, off = self._columns(s, e), self._ev_offset()
        b = dict(c)
        if off:
            b["ev"] = c["ev"] + off
        return b

    def _port_capacity(self, B):
        return self.st.max_port * B + 1 if self.procedural else super()._port_capacity(B)

    def _fill_static(self, sg, s, e):
        if not self.procedural:
            return super()._fill_static(sg, s, e)
        c, x = self._columns(s, e), sg.static
        for k in ("src", "dst", "ts", "eidx", "day", "port_ptr", "port_items"):
            x[k].copy_(c[k])
        torch.add(c["ev"], self._ev_offset(), out=x["ev"])

    def split_ranges(self):
        if not self.procedural:
            return super().split_ranges()
        return (0, self.n_train), (self.n_train, self._n_val_end), (self._n_val_end, self.st.n_events)

    def _build_finders(self, tr):
        st, tc, dev = self.st, self.tc, self.device
        uniform = tc.model == "tgat"
        if self.procedural:
            csr_tr = local_csr_from_device_stream(st, self.n_train, self.rank, self.world)
            csr_full = local_csr_from_device_stream(st, st.n_events, self.rank, self.world)
            self.csr_train, self.csr_full = csr_tr, csr_full
            self.nf_train = ShardedNeighborFinder(csr_tr, self.ex, uniform=uniform, seed=tc.seed, tag="nf_train")
            self.nf_full = ShardedNeighborFinder(csr_full, self.ex, uniform=uniform, seed=tc.seed, tag="nf_full")
            return
        csr_tr = local_csr(st.sources[tr], st.destinations[tr], st.edge_idxs[tr], st.timestamps[tr], st.n_nodes,
                           self.rank, self.world, dev)
        csr_full = local_csr(st.sources, st.destinations, st.edge_idxs, st.timestamps, st.n_nodes,
                             self.rank, self.world, dev)
        self.csr_train, self.csr_full = csr_tr, csr_full
        self.nf_train = ShardedNeighborFinder(csr_tr, self.ex, uniform=uniform, seed=tc.seed, tag="nf_train")
        self.nf_full = ShardedNeighborFinder(csr_full, self.ex, uniform=uniform, seed=tc.seed, tag="nf_full")

    def _tgn_extra(self):
        return {"memory_nodes": (self.st.n_nodes + self.world - 1) // self.world}

    def _bind_engine(self):
        tgn = self.tgn
        state = tgn.memory.state if tgn.use_memory else None
        self.engine = ShardedEngine(tgn._cfg, state, tgn.node_raw_features, tgn.edge_raw_features, self.nf_train,
                                    self.ex, self.st.n_nodes)
        tgn._engine = self.engine

    # ---- gradient bucket (same scheme as the replicated trainer)
    def _zero_grads(self):
        if self._own_bucket:
            self.gflat.zero_()

    def _reduce_grads(self):
        if self.world > 1:
            allreduce_sum_(self.gflat, self.group)

    def _all_ranks_sum(self, t):
        return allreduce_sum_(t, self.group) if self.world > 1 else t

    def _check_fit(self, bs):
        if bs < self.world:
            raise ValueError(f"fit: batch size {bs} is smaller than the number of ranks ({self.world})")

    # ---- steps
    def _fwd_bwd(self, b):
        self.ex.begin_step()                                  # inside the captured region: step-start barrier of the arena
        return super()._fwd_bwd(b)

    def _eval_body(self, b, N, with_state):
        self.ex.begin_step()
        return super()._eval_body(b, N, with_state)

    def _slice(self, s, e):
        ls, le = replica_slice(s, e, self.rank, self.world)
        self.engine.key_base, self.engine.key_side = ls - s, e - s
        self.ex.batch = (le - ls, -(-(e - s) // self.world))
        return ls, le

    def _run_graphed(self, sg):
        if sg.graph is None and sg.eager_steps >= 2:
            self.ex.freeze()                                  # capacities fixed from the eager steps' bucket counts
        eager = sg.graph is None and sg.eager_steps < 2
        out = super()._run_graphed(sg)
        if eager:
            self.ex.collect()
        return out

    def train_step(self, s, e, batch=None):
        ls, le = self._slice(s, e)
        self.engine.tag = "train"
        if (e - s) % self.world != 0 or not (self._graph_ok(le - ls) and e <= self.st.n_events):
            # ragged tail batch (or graphs off): launched kernel by kernel, this rank's mean weighted by its share
            keep, self._loss_scale = self._loss_scale, (le - ls) / float(e - s)
            try:
                out = self._step_body(self._batch(ls, le))
            finally:
                self._loss_scale = keep
            self.ex.collect()
            return out
        sg = self._step_graph(le - ls)
        self._fill_static(sg, ls, le)
        return self._run_graphed(sg)

    def eval_step(self, s, e, n_items=None, batch=None, state_batch=None):
        """Global evaluation batch [s, e): this rank scores its slice of the users (returns its slice of the results)."""
        ls, le = self._slice(s, e)
        self.engine.tag = "eval"
        N = int(n_items) if n_items is not None else int(self.eval_sampler.items.shape[0])
        sg = self._graphs.get(le - ls)
        eg = sg.evals.get((N, False)) if sg is not None else None
        replay = eg is not None and eg.graph is not None
        capture = eg is not None and eg.graph is None and eg.eager_steps >= 2 and self._graph_ok(le - ls)
        if capture:
            self.ex.freeze()
        out = super().eval_step(ls, le, n_items=n_items, ev_base=ls - s)
        if not replay and not capture:
            self.ex.collect()                                 # an eager step ran: record its bucket counts
        return out

    def make_host_batches(self, start, count, bs):
        """Pinned host copies of this rank's slices of `count` consecutive GLOBAL batches of bs * world interactions."""
        out = []
        for i in range(count):
            s = start + i * bs * self.world
            ls, le = replica_slice(s, s + bs * self.world, self.rank, self.world)
            if self.procedural:
                c, off = self._columns(ls, le), self._ev_offset()
                hb = self._pack_host_batch({k: (v + off if k == "ev" else v).cpu() for k, v in c.items()})
            else:
                hb = super().make_host_batches(ls, 1, bs)[0]
            hb["_global"] = (s, s + bs * self.world)
            out.append(hb)
        return out

    def train_step_host(self, hb):
        s, e = hb.pop("_global")
        self._slice(s, e)
        self.engine.tag = "train"
        try:
            B = hb["src"].shape[0]
            sg = self._step_graph(B)
            self._copy_host_batch(sg, hb)
            return self._run_graphed(sg)
        finally:
            hb["_global"] = (s, e)

    def gather_memory(self):
        """Global [N, .] memory / last_update / pend_valid assembled on every rank (tests)."""
        st, G = self.engine.state, self.world
        outs = []
        for t in (st.memory, st.last_update, st.pend_valid.to(torch.float32)):
            parts = [torch.empty_like(t) for _ in range(G)]
            if G > 1:
                dist.all_gather(parts, t.contiguous(), group=self.group)
            else:
                parts = [t]
            full = torch.stack(parts, dim=1).reshape((-1,) + tuple(t.shape[1:]))[:self.st.n_nodes]
            outs.append(full)
        return outs

"""Drive the UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) on the host cores.

This is the reference arm of bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`): the reference's own
`TGN`, `NeighborFinder`, `RandEdgeSampler`, `compute_time_statistics` and `eval_recommendation` are imported as they
are, and the training step is the TEXT of reference main.py:167-394 (the body of the batch loop: candidate
sampling, the inline MV-selection block, embeddings, BPR, backward, Adam step, memory detach) sliced out of the
staged main.py at run time and exec'd -- nothing of it is restated here.  What this file adds is only what
main.py:86-124 does around that body (load data, build finders / model / optimiser), fed from the synthetic
stream in memory instead of from data/period_*/ files, with `device = cpu` (main.py:103 hard-codes cuda).

Must run in a process that has NOT imported the drop-in overlay: both use the module paths model.*, modules.*,
utils.* (bench.py launches it as a subprocess from the GPU arm).
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
import textwrap
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from stage_reference import ref_root   # noqa: E402

# main.py lines (1-based, inclusive) of the batch-loop body, and what they must start / end with
BODY_FIRST, BODY_LAST = 167, 394
BODY_HEAD, BODY_TAIL = "loss = 0", "tgn.memory.detach_memory()"


def _loop_body_source(root):
    lines = open(os.path.join(root, "main.py"), encoding="utf-8").read().split("\n")
    body = lines[BODY_FIRST - 1:BODY_LAST]
    if body[0].strip() != BODY_HEAD or body[-1].strip() != BODY_TAIL:
        raise RuntimeError("staged main.py does not match the reference revision this harness slices")
    return textwrap.dedent("\n".join(body))


class ReferenceRunner:
    """The reference model + optimiser on CPU over a synthetic `Stream` (pfotgnrec_b200.synth)."""

    def __init__(self, st, model_name="ours", bs=512, n_layers=1, n_neighbors=10, dropout=0.1, lr=1e-4,
                 num_negatives=20, train_hi=None, threads=None, quiet=True):
        import torch
        root = ref_root()
        if root is None:
            raise RuntimeError("no reference tree: neither /root/reference nor baseline/_ref/main.py exists")
        self.root = root
        if root not in sys.path:
            sys.path.insert(0, root)
        if "model.tgn" in sys.modules and not sys.modules["model.tgn"].__file__.startswith(root):
            raise RuntimeError("the drop-in overlay is imported in this process: run the reference arm in its own process")
        from model.tgn import TGN
        from utils.data import Data, compute_time_statistics
        from utils.utils import RandEdgeSampler, get_neighbor_finder
        import scipy.stats as stats
        if threads:
            torch.set_num_threads(int(threads))
        torch.manual_seed(0)                                           # main.py:8-9
        np.random.seed(0)
        self.st, self.torch = st, torch
        E = st.n_events
        train_mask = st.split()[0]
        n_train = int(train_mask.sum())
        hi = n_train if train_hi is None else min(int(train_hi), n_train)
        self.n_train = hi
        # portfolios: array of lists of stock codes, [''] when empty (utils/preprocess_data.py:40-45) -- built lazily
        # for the interactions a step touches (5 M Python lists are not needed to time a bounded sample)
        self._port = np.empty(E, dtype=object)
        self._port_done = np.zeros(E, dtype=bool)
        labels = np.zeros(E, dtype=object)
        mk = lambda a, b: Data(st.sources[a:b], st.destinations[a:b], st.timestamps[a:b], st.edge_idxs[a:b],
                               labels[a:b], self._port[a:b])
        self.full_data = mk(0, E)
        self.train_data = mk(0, hi)                                    # utils/data.py:48-53 (time-sorted stream)
        self.upper_u = int(st.sources.max())
        self.map_item_id = {c: k for k, c in enumerate(st.codes)}
        self.time_feature = {dk: {c: st.prices_future[di, k] for k, c in enumerate(st.codes)}
                             for di, dk in enumerate(st.day_keys)}
        node_features = np.random.rand(st.n_nodes, 64)                # main.py:87
        t0 = time.perf_counter()
        self.train_ngh_finder = get_neighbor_finder(self.train_data, uniform=(model_name == "tgat"),
                                                    max_node_idx=st.n_nodes - 1)
        self.full_ngh_finder = None
        stats4 = compute_time_statistics(st.sources, st.destinations, st.timestamps)
        self.init_s = time.perf_counter() - t0
        a = types.SimpleNamespace(model_name=model_name, bs=int(bs), drop_out=float(dropout), lr=float(lr),
                                  num_negatives=int(num_negatives), p_pos_num=1, p_neg_num=3, gamma=2.0,
                                  lambda_mv=0.5, n_degree=int(n_neighbors), test_run=False, period="30",
                                  memory_dim=64, n_head=2, dyrep=False, memory_updater="gru",
                                  embedding_module="graph_attention", use_destination_embedding_in_message=False)
        use_memory = True
        if model_name == "jodie":                                      # main.py:63-74
            a.memory_updater, a.embedding_module = "rnn", "time"
        elif model_name == "dyrep":
            a.memory_updater, a.use_destination_embedding_in_message, a.dyrep = "rnn", True, True
        elif model_name == "tgat":
            use_memory = False
        device = torch.device("cpu")                                   # main.py:103 is cuda:{gpu}
        tgn = TGN(neighbor_finder=self.train_ngh_finder, node_features=node_features,                  # main.py:106-121
                  edge_features=st.edge_features, device=device, n_layers=int(n_layers), n_heads=a.n_head,
                  dropout=a.drop_out, use_memory=use_memory, message_dimension=100, memory_dimension=a.memory_dim,
                  memory_update_at_start=True, embedding_module_type=a.embedding_module, message_function="identity",
                  aggregator_type="last", memory_updater_type=a.memory_updater, n_neighbors=a.n_degree,
                  mean_time_shift_src=stats4[0], std_time_shift_src=stats4[1], mean_time_shift_dst=stats4[2],
                  std_time_shift_dst=stats4[3],
                  use_destination_embedding_in_message=a.use_destination_embedding_in_message,
                  use_source_embedding_in_message=False, dyrep=a.dyrep)
        optimizer = torch.optim.Adam(tgn.parameters(), lr=a.lr)       # main.py:123-124
        tgn = tgn.to(device)
        self.args = a
        self.ns = dict(np=np, torch=torch, stats=stats, RandEdgeSampler=RandEdgeSampler, args=a, tgn=tgn,
                       optimizer=optimizer, train_data=self.train_data, upper_u=self.upper_u,
                       map_item_id=self.map_item_id, time_feature=self.time_feature, BACKPROP_EVERY=1,
                       USE_MEMORY=use_memory, losses_batch=[], num_instance=hi, num_batch=0, batch=0)
        self._body = compile(_loop_body_source(root), os.path.join(root, "main.py") + ":167-394", "exec")
        tgn.set_neighbor_finder(self.train_ngh_finder)                 # main.py:156
        if not quiet:
            print(f"[ref_harness] reference at {root}; finder + time statistics {self.init_s:.1f}s", file=sys.stderr)

    def _fill_portfolios(self, s, e):
        st, todo = self.st, np.nonzero(~self._port_done[s:e])[0] + s
        for ev in todo:
            self._port[ev] = [st.codes[k] for k in st.portfolio(ev)] or [""]
        self._port_done[s:e] = True

    def train_step(self, pos, n):
        """Interactions [pos, pos + n) as ONE batch of the reference loop (pos must be a multiple of n: the body
        addresses batches as batch_idx * args.bs, main.py:179-180).  Returns the loss."""
        if pos % n != 0 or pos + n > self.n_train:
            raise ValueError("batch must be aligned to its size and inside the training split")
        self._fill_portfolios(pos, pos + n)
        ns = self.ns
        ns["args"].bs = int(n)
        ns["batch"] = pos // n
        ns["num_batch"] = -(-self.n_train // n)
        exec(self._body, ns)
        return ns["losses_batch"][-1]

    def reset_memory(self):
        if self.ns["USE_MEMORY"]:
            self.ns["tgn"].memory.__init_memory__()                    # main.py:152-153

    def evaluate(self, s, e, batch_size):
        """reference evaluation.py:39-258 over interactions [s, e) of the full stream (full-graph neighbour finder,
        all-stock ranking, metric block).  eval_recommendation reads its price pickles from ./data/period_{p}/, so
        they are written to a temporary directory first.  Returns (dict, users scored, seconds in eval_recommendation)."""
        from evaluation import eval_recommendation
        from utils.data import Data
        from utils.utils import get_neighbor_finder
        st = self.st
        if self.full_ngh_finder is None:
            self.full_ngh_finder = get_neighbor_finder(self.full_data, uniform=(self.args.model_name == "tgat"),
                                                       max_node_idx=st.n_nodes - 1)
        self._fill_portfolios(s, e)
        labels = np.zeros(e - s, dtype=object)
        data = Data(st.sources[s:e], st.destinations[s:e], st.timestamps[s:e], st.edge_idxs[s:e], labels, self._port[s:e])
        work = tempfile.mkdtemp(prefix="pfo_ref_eval_")
        d = os.path.join(work, "data", "period_30")
        os.makedirs(d)
        pickle.dump(self.map_item_id, open(os.path.join(d, "map_item_id.pkl"), "wb"))
        pickle.dump(self.time_feature, open(os.path.join(d, "time_feature_future_30.pkl"), "wb"))
        past = {dk: {c: st.prices_past[di, k] for k, c in enumerate(st.codes)} for di, dk in enumerate(st.day_keys)}
        pickle.dump(past, open(os.path.join(d, "time_feature_past_30.pkl"), "wb"))
        tgn = self.ns["tgn"]
        tgn.set_neighbor_finder(self.full_ngh_finder)                  # main.py:405
        cwd = os.getcwd()
        os.chdir(work)
        try:
            t0 = time.perf_counter()
            out = eval_recommendation(tgn=tgn, data=data, full_data=self.full_data, batch_size=int(batch_size),
                                      n_neighbors=self.args.n_degree, upper_u=self.upper_u, period="30",
                                      is_test_run=False, EVAL="test")
            dt = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
            tgn.set_neighbor_finder(self.train_ngh_finder)
        n_batches = -(-(e - s) // batch_size)
        users = sum(min(e - s, (b + 1) * batch_size) - b * batch_size for b in range(n_batches)
                    if min(e - s, (b + 1) * batch_size) != e - s)     # evaluation.py:68-69 skips the last batch
        return out, users, dt

"""Drop-in for reference modules/memory.py over the dense device state (TGNState)."""
from collections import defaultdict

import torch
from torch import nn

from pfotgnrec_b200.engine import TGNState, ModelConfig


class Memory(nn.Module):
    def __init__(self, n_nodes, memory_dimension, input_dimension, message_dimension=None,
                 device="cpu", combination_method='sum', n_edge_features=None):
        super(Memory, self).__init__()
        self.n_nodes = n_nodes
        self.memory_dimension = memory_dimension
        self.input_dimension = input_dimension
        self.message_dimension = message_dimension
        self.device = device
        self.combination_method = combination_method
        F = n_edge_features if n_edge_features is not None else input_dimension - 3 * memory_dimension
        self._state = TGNState(n_nodes, ModelConfig(d=memory_dimension, n_edge_feat=F), device)
        self.__init_memory__()

    def __init_memory__(self):
        """Zero the memory; called at the start of each epoch (reference memory.py:23-33)."""
        self._state.reset()
        # parameters (no grad) so that they are saved with the model, sharing the state's storage
        self.memory = nn.Parameter(self._state.memory, requires_grad=False)
        self.last_update = nn.Parameter(self._state.last_update, requires_grad=False)

    @property
    def state(self):
        # .to(device) may have re-homed the parameters; keep the engine's view in sync
        if self.memory.data_ptr() != self._state.memory.data_ptr():
            self._state.memory = self.memory.data
        if self.last_update.data_ptr() != self._state.last_update.data_ptr():
            self._state.last_update = self.last_update.data
        return self._state

    @property
    def messages(self):
        """Reference-shaped view {node: [(message, timestamp)]} of the pending table (debug/compat)."""
        st = self._state
        out = defaultdict(list)
        raw = st.cfg.raw
        for node in torch.nonzero(st.pend_valid).flatten().tolist():
            out[node] = [(st.pend_msg[node, :raw].clone(), st.pend_ts[node].clone())]
        return out

    def store_raw_messages(self, nodes, node_id_to_messages):
        st = self._state
        for node in nodes:
            for msg, ts in node_id_to_messages[node]:
                st.pend_msg[node, :st.cfg.raw] = msg
                st.pend_ts[node] = ts
                st.pend_valid[node] = 1

    def get_memory(self, node_idxs):
        return self.memory[node_idxs, :]

    def set_memory(self, node_idxs, values):
        self.memory[node_idxs, :] = values

    def get_last_update(self, node_idxs):
        return self.last_update[node_idxs]

    def backup_memory(self):
        b = self.state.backup()
        return b[0], b[1], b[2:]

    def restore_memory(self, memory_backup):
        self.state.restore((memory_backup[0], memory_backup[1]) + tuple(memory_backup[2]))

    def detach_memory(self):
        """The dense state never carries autograd history: nothing to detach (memory.py:62-71)."""
        return None

    def clear_messages(self, nodes):
        self._state.pend_valid[torch.as_tensor(list(nodes), dtype=torch.long, device=self._state.device)] = 0

"""Drop-in for reference modules/message_function.py: parameter containers with the reference's names and init
order; the MLP (Linear -> ReLU -> Linear on the aggregated raw message, model/tgn.py:342-354) runs inside the lazy
memory update of the step engine (pfo_linear_* / pfo_wgrad_*)."""
from torch import nn


class MessageFunction(nn.Module):
    def compute_message(self, raw_messages):
        return None


class MLPMessageFunction(MessageFunction):
    def __init__(self, raw_message_dimension, message_dimension):
        super(MLPMessageFunction, self).__init__()
        self.mlp = self.layers = nn.Sequential(
            nn.Linear(raw_message_dimension, raw_message_dimension // 2), nn.ReLU(),
            nn.Linear(raw_message_dimension // 2, message_dimension))

    def compute_message(self, raw_messages):
        raise NotImplementedError("the message MLP is evaluated inside TGN.compute_temporal_embeddings*")


class IdentityMessageFunction(MessageFunction):
    def compute_message(self, raw_messages):
        return raw_messages


def get_message_function(module_type, raw_message_dimension, message_dimension):
    if module_type == "mlp":
        return MLPMessageFunction(raw_message_dimension, message_dimension)
    elif module_type == "identity":
        return IdentityMessageFunction()
    raise ValueError(module_type)

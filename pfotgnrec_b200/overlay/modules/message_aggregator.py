"""Drop-in for reference modules/message_aggregator.py.  The `last` aggregator is the last-wins
scatter of pfo_store_messages, `mean` the in-order segment mean of pfo_store_messages_mean (dense pending
table either way); the classes remain as the API surface over the dict-of-lists compat view."""
import torch


class MessageAggregator(torch.nn.Module):
    def __init__(self, device):
        super(MessageAggregator, self).__init__()
        self.device = device

    def aggregate(self, node_ids, messages):
        raise NotImplementedError


class LastMessageAggregator(MessageAggregator):
    def aggregate(self, node_ids, messages):
        """Dict-of-lists view (compat): keep the last message of every node that has one."""
        ids, msgs, tss = [], [], []
        for node_id in sorted(set(int(x) for x in node_ids)):
            if len(messages[node_id]) > 0:
                ids.append(node_id)
                msgs.append(messages[node_id][-1][0])
                tss.append(messages[node_id][-1][1])
        return ids, (torch.stack(msgs) if ids else []), (torch.stack(tss) if ids else [])


class MeanMessageAggregator(MessageAggregator):
    def aggregate(self, node_ids, messages):
        """Dict-of-lists view (compat): the dense table already holds the mean of every node's list."""
        ids, msgs, tss = [], [], []
        for node_id in sorted(set(int(x) for x in node_ids)):
            if len(messages[node_id]) > 0:
                ids.append(node_id)
                msgs.append(torch.mean(torch.stack([m[0] for m in messages[node_id]]), dim=0))
                tss.append(messages[node_id][-1][1])
        return ids, (torch.stack(msgs) if ids else []), (torch.stack(tss) if ids else [])


def get_message_aggregator(aggregator_type, device):
    if aggregator_type == "last":
        return LastMessageAggregator(device=device)
    elif aggregator_type == "mean":
        return MeanMessageAggregator(device=device)
    raise ValueError("Message aggregator {} not implemented".format(aggregator_type))

"""Module path of reference modules/message_aggregator.py in the drop-in overlay: re-exports the parameter containers of
pfotgnrec_b200/containers.py (the arithmetic runs in libpfo_b200.so behind TGN.compute_temporal_embeddings*)."""
from pfotgnrec_b200.containers import (  # noqa: F401
    MessageAggregator,
    LastMessageAggregator,
    MeanMessageAggregator,
    get_message_aggregator,
)

"""relationprediction_b200 -- B200-native R-GCN relational message-passing hot path.

Host-side mirror of the reference's plugin surface (model.Model chain, common.model_builder,
encoders.*, decoders.*, extras.graph_representations) over the C-ABI library librgcn_b200.so."""
__version__ = "0.1.0"

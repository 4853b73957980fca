"""RelationEmbedding (reference: encoders/relation_embedding.py:5-25): owns the DistMult relation table and
splices it into the (subject codes, relation codes, object codes) triple coming up the chain.

The table's shape is whatever the factory passes -- for the R-GCN encoders that is [EntityCount,
CodeDimension], not [RelationCount, CodeDimension] (quirk Q4, model_builder.py:134-135): rows beyond
RelationCount exist, are initialised, are listed in get_weights() and are simply never gathered.  The scorer
kernels index the table with the relation id, so the oversize table costs memory, not time."""
import numpy as np
import torch

from ..model import Model


class RelationEmbedding(Model):
    shape = None

    def __init__(self, shape, settings, next_component=None):
        Model.__init__(self, next_component, settings)
        self.shape = shape

    @property
    def embedding_width(self):
        """CodeDimension of the settings (the reference caches it in parse_settings, :12-13)."""
        return int(self.settings['CodeDimension'])

    def local_initialize_train(self):
        # standard-normal rows drawn from numpy's global stream, as the reference does (:16), then uploaded
        rows, width = int(self.shape[0]), int(self.shape[1])
        table = torch.from_numpy(np.random.randn(rows, width).astype(np.float32)).to(self.get_device())
        self.W_relation = table.requires_grad_(True)

    def local_get_weights(self):
        return [self.W_relation]

    def get_all_codes(self, mode='train'):
        subject_codes, _, object_codes = self.next_component.get_all_codes(mode=mode)
        return subject_codes, self.W_relation, object_codes

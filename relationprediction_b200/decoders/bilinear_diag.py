"""BilinearDiag = DistMult decoder (reference: decoders/bilinear_diag.py)."""
import numpy as np
import torch

from ..model import Model, Placeholder
from .. import ops


class BilinearDiag(Model):
    def __init__(self, next_component, settings):
        self.encoder_cache = {'train': None, 'test': None}
        self._scored = {'train': None, 'test': None}
        Model.__init__(self, next_component, settings)

    def parse_settings(self):
        self.regularization_parameter = float(self.settings['RegularizationParameter'])

    def local_initialize_train(self):
        self.Y = Placeholder('Y', 'float32', [None])
        self.X = Placeholder('X', 'int32', [None, 3])

    def local_clear_cache(self):
        self.encoder_cache = {'train': None, 'test': None}
        self._scored = {'train': None, 'test': None}

    def local_get_train_input_variables(self):
        return [self.X, self.Y]

    def local_get_test_input_variables(self):
        return [self.X]

    def _x_device(self):
        dev = self.get_device()
        return torch.as_tensor(np.ascontiguousarray(np.asarray(self.X.value, dtype=np.int32).reshape(-1, 3)),
                               device=dev)

    def _fused(self, mode):
        """One fused kernel: three row gathers -> energies (+ sigmoid-CE loss and L2 term in train
        mode).  The gathered e1/r/e2 rows of compute_codes are never materialised."""
        if self._scored[mode] is None:
            subject_codes, relation_codes, object_codes = self.next_component.get_all_codes(mode=mode)
            assert subject_codes is object_codes, "DistMult kernel expects one shared entity code matrix"
            Y = None
            if mode == 'train':
                Y = torch.as_tensor(np.asarray(self.Y.value, dtype=np.float32), device=self.get_device())
            self._scored[mode] = ops.distmult(subject_codes.contiguous(), relation_codes.contiguous(),
                                              self._x_device(), Y)
        return self._scored[mode]

    def compute_codes(self, mode='train'):
        """(e1s, rs, e2s) row gathers (bilinear_diag.py:14-24) -- only the all-entity scoring GEMMs
        below need them explicitly."""
        if self.encoder_cache[mode] is None:
            subject_codes, relation_codes, object_codes = self.next_component.get_all_codes(mode=mode)
            X = self._x_device().long()
            self.encoder_cache[mode] = (subject_codes[X[:, 0]], relation_codes[X[:, 1]], object_codes[X[:, 2]])
        return self.encoder_cache[mode]

    def get_loss(self, mode='train'):
        return self._fused(mode)[1]  # reduce_mean(weighted CE, pos_weight forced to 1) (:27-34)

    def local_get_regularization(self):
        return self.regularization_parameter * self._fused('train')[2]  # (:63-69)

    def predict(self):
        return torch.sigmoid(self._fused('test')[0])

    def predict_all_subject_scores(self):
        e1s, rs, e2s = self.compute_codes(mode='test')
        all_subject_codes = self.next_component.get_all_subject_codes(mode='test')
        return torch.sigmoid((all_subject_codes @ (rs * e2s).T).T)

    def predict_all_object_scores(self):
        e1s, rs, e2s = self.compute_codes(mode='test')
        all_object_codes = self.next_component.get_all_object_codes(mode='test')
        return torch.sigmoid((e1s * rs) @ all_object_codes.T)

    # ---- fused all-entity scoring + ranking (next row N3; library entry distmult_rank) ----
    @staticmethod
    def known_bit_mask(lists, n_entities):
        """int32 view of the uint32 [n, ceil(V/32)] bit masks the library expects: bit v of row t = v in lists[t]."""
        words = (n_entities + 31) // 32
        m = np.zeros((len(lists), words), np.uint32)
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        if lens.sum():
            rows = np.repeat(np.arange(len(lists)), lens)
            cols = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists if len(l)])
            np.bitwise_or.at(m, (rows, cols >> 5), np.left_shift(np.uint32(1), (cols & 31).astype(np.uint32)))
        return m.view(np.int32)

    def rank_all(self, triplets, known_subject_lists, known_object_lists, chunk=4096):
        """Raw and filtered ranks of every triple under subject and object corruption with the rules of
        common/evaluation.py:148-159 / :355-367, without materialising the [n, V] score matrices of
        predict_all_subject_scores / predict_all_object_scores (bilinear_diag.py:51-61): the encoder runs once,
        the scoring GEMM counts `score >= gold` in its epilogue.  Returns four int arrays
        (raw_subjects, filtered_subjects, raw_objects, filtered_objects)."""
        subject_codes, relation_codes, object_codes = self.next_component.get_all_codes(mode='test')
        assert subject_codes is object_codes, "DistMult ranking expects one shared entity code matrix"
        codes, rel = subject_codes.contiguous(), relation_codes.contiguous()
        ranker = ops.DistMultRanker(codes, rel)
        V = codes.shape[0]
        tri = np.ascontiguousarray(np.asarray(triplets, dtype=np.int32).reshape(-1, 3))
        out = [[], [], [], []]
        for c0 in range(0, len(tri), chunk):
            X = torch.as_tensor(tri[c0:c0 + chunk], device=codes.device)
            for side, lists in ((0, known_subject_lists), (1, known_object_lists)):
                mask = torch.as_tensor(self.known_bit_mask(lists[c0:c0 + chunk], V), device=codes.device)
                raw, filt = ranker.rank(X, side, mask)
                out[2 * side].append(raw)
                out[2 * side + 1].append(filt)
        return tuple(torch.cat(o).cpu().numpy().astype(np.int64) if o else np.zeros(0, np.int64) for o in out)

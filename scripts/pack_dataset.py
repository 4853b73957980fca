"""Pack a dataset directory (train/valid/test.txt + entities.dict + relations.dict, the reference's format)
into one .npz for `python -m relationprediction_b200.train --dataset-npz`.  Run HERE (the datasets live in the
reference tree); the output goes to the untracked .scratch/ directory, which travels to the GPU box.

  python scripts/pack_dataset.py /root/reference/data/FB-Toutanova .scratch/fb15k237_full.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relationprediction_b200.train import load_dataset  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
splits, entities, relations = load_dataset(src)
os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
np.savez_compressed(dst, V=len(entities), R=len(relations),
                    **{k: np.asarray(v, dtype=np.int32) for k, v in splits.items()})
print(dst, {k: len(v) for k, v in splits.items()}, len(entities), len(relations))

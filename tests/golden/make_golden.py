"""Generates the committed golden fixtures.  Run HERE (needs /root/reference); the outputs travel.

  python tests/golden/make_golden.py

* toy_golden.json   -- data/Toy as integer triples, loaded with the REFERENCE's own numpy-only loader
                       (common/io.py:27-39), the SURVEY.md 8(c) integer goldens (degrees, norms), the
                       raw text of the shipped .exp settings files and their parse by the REFERENCE's
                       settings_reader (common/settings_reader.py:29-48).
* (reference-code goldens: see make_reference_golden.py)
* layer_golden.npz  -- seeded inputs + float64 oracle outputs (forward, all gradients) of one block
                       layer and one basis layer on the Toy graph, plus DistMult loss/grad.  The
                       reference itself cannot run (TensorFlow 1.4 absent): these pin the ORACLE
                       restatement so that a later edit of oracle/ cannot silently drift.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REF, "code"))
sys.path.insert(0, ROOT)

from common import io as ref_io  # noqa: E402  (reference module, numpy only)
from common import settings_reader as ref_settings  # noqa: E402
from oracle import rgcn_oracle as oracle  # noqa: E402


def settings_to_dict(s):
    return {k: (settings_to_dict(s[k]) if isinstance(s[k], ref_settings.Settings) else s[k]) for k in s}


def main():
    toy = os.path.join(REF, "data", "Toy")
    ent, rel = os.path.join(toy, "entities.dict"), os.path.join(toy, "relations.dict")
    out = {}
    for split in ("train", "valid", "test"):
        out[split] = ref_io.read_triplets_as_list(os.path.join(toy, split + ".txt"), ent, rel)
    out["entities"] = {str(k): v for k, v in ref_io.read_dictionary(ent).items()}
    out["relations"] = {str(k): v for k, v in ref_io.read_dictionary(rel).items()}
    tr = np.array(out["train"])
    V, R = len(out["entities"]), len(out["relations"])
    out["V"], out["R"] = V, R
    out["indeg_forward_by_object"] = np.bincount(tr[:, 2], minlength=V).tolist()
    out["indeg_backward_by_subject"] = np.bincount(tr[:, 0], minlength=V).tolist()
    nf, nb = oracle.graph_norms(tr, V, "canonical")
    cf, cb = oracle.graph_norms(tr, V, "tf_unsorted_compat")
    out["norm_f_canonical"] = [float(x) for x in nf]
    out["norm_b_canonical"] = [float(x) for x in nb]
    out["norm_f_tf_unsorted_compat"] = [float(x) for x in cf]
    out["norm_b_tf_unsorted_compat"] = [float(x) for x in cb]
    out["settings_text"] = {}
    out["settings_parsed"] = {}
    for name in ("gcn_block.exp", "gcn_basis.exp", "distmult.exp", "complex.exp"):
        p = os.path.join(REF, "settings", name)
        with open(p) as fh:
            out["settings_text"][name] = fh.read()
        out["settings_parsed"][name] = settings_to_dict(ref_settings.read(p))
    # dataset-level integer goldens of the other configs
    stats = {}
    for ds in ("FB-Toutanova", "wn18", "FB15k"):
        d = os.path.join(REF, "data", ds)
        e, r = os.path.join(d, "entities.dict"), os.path.join(d, "relations.dict")
        t = np.array(ref_io.read_triplets_as_list(os.path.join(d, "train.txt"), e, r))
        nv, nr = len(ref_io.read_dictionary(e)), len(ref_io.read_dictionary(r))
        deg = np.bincount(np.concatenate([t[:, 0], t[:, 2]]), minlength=nv)
        stats[ds] = {"V": nv, "R": nr, "E_train": int(t.shape[0]), "isolated": int((deg == 0).sum()),
                     "max_degree": int(deg.max()),
                     "first_triples": t[:5].tolist(),
                     "checksum_s_r_o": [int(t[:, 0].sum()), int(t[:, 1].sum()), int(t[:, 2].sum())]}
    out["dataset_stats"] = stats
    with open(os.path.join(HERE, "toy_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)

    # ---- numeric goldens (float64 oracle) on the Toy graph ----
    rng = np.random.RandomState(0)
    d, B_block, B_basis = 8, 2, 2
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    mask = (rng.uniform(size=(V, d)) < 0.8).astype(np.uint8)
    arrs = {"H": H, "dOut": dOut, "mask": mask, "triples": tr.astype(np.int32)}
    wb = oracle.init_block_layer(rng, R, d, B_block)
    ws = oracle.init_basis_layer(rng, R, d, B_basis)
    for variant, w in (("block", wb), ("basis", ws)):
        for k, v in w.items():
            arrs["%s_%s" % (variant, k)] = v
        for tag, m, keep, relu in (("plain", None, 1.0, True), ("drop", mask, 0.8, False)):
            o, g = oracle.layer_fwd_bwd(variant, H, tr, w, nf, nb, dOut, m, keep, relu, torch.float64)
            arrs["%s_%s_out" % (variant, tag)] = o.numpy()
            for k, v in g.items():
                arrs["%s_%s_d%s" % (variant, tag, k)] = v.numpy()
    # DistMult
    N = 24
    X = np.stack([rng.randint(0, V, N), rng.randint(0, R, N), rng.randint(0, V, N)], 1).astype(np.int32)
    Y = (rng.uniform(size=N) < 0.3).astype(np.float32)
    codes = rng.normal(0, 1, (V, d)).astype(np.float32)
    relt = rng.normal(0, 1, (V, d)).astype(np.float32)  # [EntityCount, d] (quirk Q4)
    ct = torch.tensor(codes, dtype=torch.float64, requires_grad=True)
    rt = torch.tensor(relt, dtype=torch.float64, requires_grad=True)
    loss, reg, en = oracle.distmult_loss(ct, rt, X, Y, torch.float64)
    (loss + 0.01 * reg).backward()
    arrs.update({"dm_X": X, "dm_Y": Y, "dm_codes": codes, "dm_rel": relt, "dm_energies": en.detach().numpy(),
                 "dm_loss": np.array(loss.item()), "dm_reg": np.array(reg.item()),
                 "dm_dcodes": ct.grad.numpy(), "dm_drel": rt.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "layer_golden.npz"), **arrs)
    print("wrote toy_golden.json, layer_golden.npz")


if __name__ == "__main__":
    main()


def make_eval_golden():
    """Ranking golden from the REFERENCE's own numpy-only evaluation.py (importable): a fake model with
    tied scores, raw + filtered MRR / Hits@n over both corruption directions."""
    from common import evaluation as ref_eval
    rng = np.random.RandomState(0)
    V = 50
    train = np.stack([rng.randint(0, V, 300), rng.randint(0, 4, 300), rng.randint(0, V, 300)], 1)
    T = np.round(rng.rand(4, V, V).astype(np.float32), 2)

    class FakeModel:
        def score_all_subjects(self, tr):
            return np.stack([T[r, :, o] for s, r, o in tr])

        def score_all_objects(self, tr):
            return np.stack([T[r, s, :] for s, r, o in tr])
    sc = ref_eval.Scorer({'Metric': 'MRR'})
    sc.register_data(train)
    sc.register_degrees(train)
    sc.register_model(FakeModel())
    sc.finalize_frequency_computation(train)
    summ = sc.compute_scores(train[:40]).get_summary()
    res = {k: {m: float(v) for m, v in summ.results[k].items() if m in ('MRR', 'H@1', 'H@3', 'H@10')}
           for k in ('Raw', 'Filtered')}
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), train=train, T=T,
                        raw=np.array([res['Raw'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')]),
                        filtered=np.array([res['Filtered'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')]))
    # negative sampler golden (reference auxilliaries.NegativeSampler, numpy-only, seeded)
    from common import auxilliaries as ref_aux
    np.random.seed(123)
    ns = ref_aux.NegativeSampler(3, V)
    idx, lab = ns.transform(train[:20])
    np.savez_compressed(os.path.join(HERE, "negsample_golden.npz"), batch=train[:20], idx=idx, labels=lab)


if __name__ == "__main__":
    make_eval_golden()
    print("wrote eval_golden.npz, negsample_golden.npz")

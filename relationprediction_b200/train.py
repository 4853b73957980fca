"""Training driver mirroring the reference's code/train.py flow (next rows N1-N3):

  python -m relationprediction_b200.train --settings X.exp --dataset DIR [--max-iterations N]

settings/dataset formats, the section merge (train.py:69-86), the per-step sample transform
(graph batch -> GraphSplitSize edge dropout -> negative sampling, :201-245), loss = CE + regularisation
(:262), global-norm clipping + Adam (optimization/tensorflow_backend/algorithms.py:36-42, :65-68) and
the periodic validation MRR follow the reference; the "Converge" optimizer stack itself is replaced by
a plain loop.  Clip + Adam run on the library's kernels with the TensorFlow-1.x formulas (optim.py,
csrc/optimizer.cu)."""
import argparse
import os

import numpy as np
import torch

from .optim import ClippedAdam
from .common import auxilliaries, evaluation, io, model_builder, settings_reader


def load_dataset(dataset):
    ent, rel = os.path.join(dataset, 'entities.dict'), os.path.join(dataset, 'relations.dict')
    splits = {k: io.read_triplets_as_array(os.path.join(dataset, k + '.txt'), ent, rel)
              for k in ('train', 'valid', 'test')}
    return splits, io.read_dictionary(ent), io.read_dictionary(rel)


def merge_settings(settings, n_entities, n_relations, n_train):
    general = settings['General']
    general.put('EntityCount', n_entities)
    general.put('RelationCount', n_relations)
    general.put('EdgeCount', n_train)
    for name in ('Encoder', 'Decoder'):
        settings[name].merge(settings['Shared'])
        settings[name].merge(general)
    settings['Optimizer'].merge(general)
    settings['Evaluation'].merge(general)
    return settings


def sample_edge_neighborhood_fast(triples, n_entities, sample_size):
    """The same sampler in the library (csrc/sampler.cu): Fenwick trees instead of an O(V) np.random.choice per
    draw (~5 s -> ~10 ms for 30 000 edges of FB15k-237); seeded from numpy's global stream."""
    import ctypes

    from . import _lib
    tri = np.ascontiguousarray(triples, dtype=np.int32)
    out = np.empty(sample_size, dtype=np.int32)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    rc = _lib.load().rgcn_sample_edge_neighborhood(ctypes.c_void_p(tri.ctypes.data), tri.shape[0], int(n_entities),
                                                   int(sample_size), seed, ctypes.c_void_p(out.ctypes.data))
    _lib.check(rc, "rgcn_sample_edge_neighborhood")
    return out


def sample_edge_neighborhood(adj_list, degrees, n_triplets, sample_size):
    """Neighbourhood-expansion edge sampler (train.py:161-198), same sequential algorithm (reference
    restatement; kept as the statistical oracle of sample_edge_neighborhood_fast)."""
    edges = np.zeros(sample_size, dtype=np.int32)
    sample_counts = degrees.copy()
    picked = np.zeros(n_triplets, dtype=bool)
    seen = np.zeros(len(degrees), dtype=bool)
    for i in range(sample_size):
        weights = sample_counts * seen
        if weights.sum() == 0:
            weights = np.ones_like(weights)
            weights[sample_counts == 0] = 0
        chosen_vertex = np.random.choice(len(degrees), p=weights / weights.sum())
        chosen_adj = adj_list[chosen_vertex]
        seen[chosen_vertex] = True
        while True:
            edge_number, other = chosen_adj[np.random.randint(len(chosen_adj))]
            if not picked[edge_number]:
                break
        edges[i] = edge_number
        picked[edge_number] = True
        sample_counts[chosen_vertex] -= 1
        sample_counts[other] -= 1
        seen[other] = True
    return edges


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train a model on a given dataset.")
    ap.add_argument("--settings", required=True)
    ap.add_argument("--dataset", required=True)
    ap.add_argument("--max-iterations", type=int, default=None)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)

    settings = settings_reader.read(args.settings)
    print(settings)
    splits, entities, relations = load_dataset(args.dataset)
    train, valid, test = splits['train'], splits['valid'], splits['test']
    merge_settings(settings, len(entities), len(relations), len(train))
    general, opt = settings['General'], settings['Optimizer']

    encoder = model_builder.build_encoder(settings['Encoder'], train)
    model = model_builder.build_decoder(encoder, settings['Decoder'])
    model.set_device(args.device)
    model.preprocess(train)
    model.register_for_test(train)
    model.initialize_train()

    scorer = evaluation.Scorer(settings['Evaluation'])
    for part in (train, valid, test):
        scorer.register_data(part)
    scorer.register_model(model)

    ns = auxilliaries.NegativeSampler(int(general['NegativeSampleRate']), len(entities))
    adj_list = [[] for _ in entities]
    for i, (s, _, o) in enumerate(train.tolist()):
        adj_list[s].append((i, o))
        adj_list[o].append((i, s))
    degrees = np.array([len(a) for a in adj_list])

    def sample():
        if not encoder.needs_graph():
            X, Y = ns.transform(train)
            return (X, Y)
        if 'GraphBatchSize' in general and int(general['GraphBatchSize']) < len(train):
            ids = sample_edge_neighborhood_fast(train, len(entities), int(general['GraphBatchSize']))
        else:
            ids = np.arange(len(train))
        graph_batch = train[ids]
        split = int(float(general['GraphSplitSize']) * len(graph_batch))
        graph_split = train[np.random.choice(ids, size=split, replace=False)]
        X, Y = ns.transform(graph_batch)
        return (graph_split, X, Y)

    weights = [w for w in model.get_weights()]
    algo = opt['Algorithm']
    lr = float(algo['learning_rate'])
    max_norm = float(opt['MaxGradientNorm']) if 'MaxGradientNorm' in opt else None
    optimizer = ClippedAdam(weights, lr=lr, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=max_norm)
    report_every = int(opt['ReportTrainLossEvery']) if 'ReportTrainLossEvery' in opt else 100
    check_every = int(opt['EarlyStopping']['CheckEvery']) if 'EarlyStopping' in opt else None
    max_it = args.max_iterations if args.max_iterations is not None else 10 ** 9

    running, it = 0.0, 0
    while it < max_it:
        it += 1
        optimizer.zero_grad()
        loss = model.train_loss(*sample())
        loss.backward()
        optimizer.step()
        running += float(loss.detach())
        if it == 1:
            print("Initial loss: %f" % running)
        if it % report_every == 0:
            print("Average train loss for iteration %d-%d: %f" % (it - report_every + 1, it, running / report_every))
            running = 0.0
        if check_every and it % check_every == 0:
            summary = scorer.compute_scores(valid).get_summary()
            print("Validation filtered MRR at iteration %d: %f" % (it, summary.results['Filtered']['MRR']))
            scorer.compute_scores(test).get_summary().pretty_print()
    return model, scorer


if __name__ == "__main__":
    main()

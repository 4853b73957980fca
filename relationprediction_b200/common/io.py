"""Dataset readers for the reference's on-disk formats (reference: common/io.py:5-39).

  entities.dict / relations.dict :  <id> TAB <name>            one per line
  train.txt / valid.txt / test.txt :  <subject name> TAB <relation name> TAB <object name>

Everything funnels through one tab-splitting line reader; the integer-triple forms map names through the two
dictionaries and keep FILE ORDER (the sampler and the tf_unsorted_compat norm mode depend on it)."""
import numpy as np


def _tab_rows(path):
    with open(path, "r") as fh:
        for raw in fh:
            row = raw.strip().split("\t")
            if row != ['']:
                yield row


def read_dictionary(filename, id_lookup=True):
    """{id: name} when id_lookup else {name: id}."""
    pairs = ((int(row[0]), row[1]) for row in _tab_rows(filename) if len(row) >= 2)
    return dict(pairs) if id_lookup else {name: ident for ident, name in pairs}


def read_triplets(filename):
    """Generator over [subject, relation, object] NAME triples."""
    return _tab_rows(filename)


def read_triplet_file(filename):
    return list(_tab_rows(filename))


def read_triplets_as_array(filename, entity_dict, relation_dict):
    """int32 [n,3] array of (subject id, relation id, object id)."""
    entity_id = read_dictionary(entity_dict, id_lookup=False)
    relation_id = read_dictionary(relation_dict, id_lookup=False)
    flat = np.fromiter((ident for row in _tab_rows(filename)
                        for ident in (entity_id[row[0]], relation_id[row[1]], entity_id[row[2]])), dtype=np.int32)
    return flat.reshape(-1, 3)


def read_triplets_as_list(filename, entity_dict, relation_dict):
    """The same as nested Python lists (the reference's return type)."""
    return read_triplets_as_array(filename, entity_dict, relation_dict).tolist()

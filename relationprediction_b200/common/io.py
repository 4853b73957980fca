"""Dataset I/O in the reference's on-disk formats (common/io.py:5-39 in the reference):
dictionaries are `id<TAB>name` lines, triple files are `subject<TAB>relation<TAB>object` by NAME."""
import numpy as np


def read_dictionary(filename, id_lookup=True):
    """id -> name (id_lookup=True) or name -> id."""
    out = {}
    with open(filename, "r") as fh:
        for line in fh:
            fields = line.strip().split("\t")
            if len(fields) < 2:
                continue
            if id_lookup:
                out[int(fields[0])] = fields[1]
            else:
                out[fields[1]] = int(fields[0])
    return out


def read_triplets(filename):
    with open(filename, "r") as fh:
        for line in fh:
            yield line.strip().split("\t")


def read_triplet_file(filename):
    return list(read_triplets(filename))


def read_triplets_as_list(filename, entity_dict, relation_dict):
    """[[subject_id, relation_id, object_id], ...] in file order."""
    ent = read_dictionary(entity_dict, id_lookup=False)
    rel = read_dictionary(relation_dict, id_lookup=False)
    return [[ent[t[0]], rel[t[1]], ent[t[2]]] for t in read_triplets(filename)]


def read_triplets_as_array(filename, entity_dict, relation_dict):
    return np.asarray(read_triplets_as_list(filename, entity_dict, relation_dict), dtype=np.int32).reshape(-1, 3)

"""Work-list discovery for cross-layer equalization.

Mirror of the reference's ``utils/relation.py`` (jakc4103/DFQ): same names, arguments and results
(`Relation`, `create_relation`), pure Python graph walking - no tensor arithmetic lives here.

Reference semantics reproduced (file:line under /root/reference):
  utils/relation.py:5-27   Relation: (first, second, bn) keys + accumulated scale vector S
  utils/relation.py:33-48  walk back from a target layer through single-consumer BatchNorm2d / ReLU /
                           QuantMeasure / AvgPool2d / "F.pad" / "torch.mean" nodes to the previous
                           target layer, remembering the last BatchNorm2d seen
  utils/relation.py:50-58  consumer counts
  utils/relation.py:61-68  one relation per `first` layer; a second consumer REMOVES the entry (quirk Q4)
  utils/relation.py:70-92  delete_single: keep only chains of >= 2 relations (detection models)
"""
from collections import OrderedDict

from torch.nn import AvgPool2d, BatchNorm2d, ReLU

from .quantize import QConv2d, QuantMeasure

_PASS_THROUGH_TYPES = (BatchNorm2d, ReLU, QuantMeasure, AvgPool2d)
_PASS_THROUGH_FUNCS = ("F.pad", "torch.mean")


class Relation():
    """Two consecutive target layers whose shared channels are rescaled, and the BN between them."""

    def __init__(self, layer_idx_1, layer_idx_2, bn_idx_1):
        self.layer_first = layer_idx_1
        self.layer_second = layer_idx_2
        self.bn_idx = bn_idx_1
        self.S = None

    def __repr__(self):
        return '({}, {})'.format(self.layer_first, self.layer_second)

    def get_idxs(self):
        return self.layer_first, self.layer_second, self.bn_idx

    def set_scale_vec(self, S):
        # relation.py:20-24: the first call keeps the tensor itself, later calls multiply in place
        if self.S is None:
            self.S = S
        else:
            self.S *= S

    def get_scale_vec(self):
        return self.S


def _consumer_counts(graph, bottoms):
    counts = {}
    for key in graph:
        if key == "Data":
            continue
        for src in bottoms[key]:
            counts[src] = counts.get(src, 0) + 1
    return counts


def _is_pass_through(graph, key):
    node = graph[key]
    if type(node) in _PASS_THROUGH_TYPES:
        return True
    return type(node) == str and any(tag in key for tag in _PASS_THROUGH_FUNCS)


def _previous_target(graph, bottoms, key, targ_type, counts):
    """(previous target layer, last BN on the way) or (None, None)."""
    srcs = bottoms[key]
    last_bn = None
    while len(srcs) == 1 and srcs[0] != "Data" and counts[srcs[0]] == 1:
        cur = srcs[0]
        node = graph[cur]
        if type(node) == BatchNorm2d:
            last_bn = cur
        if type(node) in targ_type:
            return cur, last_bn
        if not _is_pass_through(graph, cur):
            return None, None
        srcs = bottoms[cur]
    return None, None


def _chains_of_two_or_more(relations):
    groups = []
    for rel in relations:
        home = -1
        for gi, group in enumerate(groups):
            if any(rel.get_idxs()[0] == other.get_idxs()[1] for other in group):
                home = gi            # the reference keeps scanning: the last matching group wins
        if home >= 0:
            groups[home].append(rel)
        else:
            groups.append([rel])
    kept = []
    for group in groups:
        if len(group) > 1:
            kept.extend(group)
    return kept


def create_relation(graph, bottoms, targ_type=[QConv2d], delete_single=False):
    counts = _consumer_counts(graph, bottoms)
    by_first = OrderedDict()
    for key in graph:
        if type(graph[key]) not in targ_type:
            continue
        prev, bn = _previous_target(graph, bottoms, key, targ_type, counts)
        if prev in by_first:
            by_first.pop(prev)
        elif prev is not None:
            by_first[prev] = Relation(prev, key, bn)
    relations = list(by_first.values())
    if delete_single:
        return _chains_of_two_or_more(relations)
    return relations

"""dfq_b200.trace: the torch.fx graph/bottoms producer against the reference tracer's committed topologies and against
hand-written expectations."""
import json
import os
import re

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from dfq_b200 import trace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _shape(graph, bottoms):
    """Structure only: [(type-or-stem, [indices of bottoms])] in graph order."""
    idx = {k: i for i, k in enumerate(graph)}
    out = []
    for k, v in graph.items():
        t = type(v).__name__ if isinstance(v, nn.Module) else re.sub(r"_?\d+$", "", v)
        out.append((t, None if bottoms[k] is None else [idx[b] for b in bottoms[k]]))
    return out


def _gold_shape(name):
    topo = json.load(open(os.path.join(GOLD, "topology_%s.json" % name)))
    idx = {n["key"]: i for i, n in enumerate(topo["nodes"])}
    out = []
    for n in topo["nodes"]:
        t = re.sub(r"_?\d+$", "", n["key"]) if n["type"] in ("Func", "Data") else n["type"]
        t = "add" if t == "iadd" else t     # fx sees `out += identity` as operator.add (Proxy has no __iadd__); the walks test 'add' in key
        out.append((t, None if n["bottoms"] is None else [idx[b] for b in n["bottoms"]]))
    return out


def test_resnet18_matches_the_reference_tracers_topology():
    tv = pytest.importorskip("torchvision")
    model = tv.models.resnet18().eval()
    graph, bottoms = trace.trace_graph(model)
    assert _shape(graph, bottoms) == _gold_shape("resnet18")
    # module nodes are keyed by id(module) and ARE the model's modules: calibrating the graph calibrates the model
    assert graph[id(model.conv1)] is model.conv1 and graph[id(model.fc)] is model.fc


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(8, 16, 3, padding=1, bias=False); self.b1 = nn.BatchNorm2d(16)
        self.c2 = nn.Conv2d(16, 16, 3, padding=1, groups=16, bias=False); self.b2 = nn.BatchNorm2d(16)
        self.c3 = nn.Conv2d(16, 8, 1, bias=False); self.b3 = nn.BatchNorm2d(8)
        self.act = nn.ReLU()
        self.side = nn.Conv2d(8, 8, 1); self.sb = nn.BatchNorm2d(8)
        self.fc = nn.Linear(16, 5)

    def forward(self, x):
        y = self.act(self.b1(self.c1(x)))
        y = F.relu(self.b2(self.c2(y)))                 # functional activation -> module node
        y = self.b3(self.c3(y))
        y = x + y                                       # residual
        z = self.sb(self.side(F.pad(x, (0, 0, 0, 0))))
        y = torch.cat([y, z], 1)
        y = y.mean(3).mean(2)
        return self.fc(y.view(y.size(0), -1))


def test_functional_ops_reuse_and_relations():
    from dfq_b200.utils.relation import create_relation
    m = _Block().eval()
    graph, bottoms = trace.trace_graph(m)
    kinds = [t for t, _ in _shape(graph, bottoms)]
    assert kinds == ["Data", "Conv2d", "BatchNorm2d", "ReLU", "Conv2d", "BatchNorm2d", "ReLU", "Conv2d", "BatchNorm2d", "add",
                     "F.pad", "Conv2d", "BatchNorm2d", "torch.cat", "torch.mean", "torch.mean", "view", "Linear"]
    keys = list(graph)
    assert bottoms[keys[9]] == ["Data", keys[8]]                        # x + y
    assert bottoms[keys[13]] == [keys[9], keys[12]]                     # cat([y, z])
    assert bottoms[keys[11]] == [keys[10]] and bottoms[keys[10]] == ["Data"]
    # the walks accept it: conv1 -> depthwise -> pointwise is one equalization chain
    for k in (id(m.b1), id(m.b2), id(m.b3), id(m.sb)):                  # what merge_batchnorm would have registered
        graph[k].fake_weight = torch.ones(graph[k].num_features); graph[k].fake_bias = torch.zeros(graph[k].num_features)
    rels = create_relation(graph, bottoms, [nn.Conv2d, nn.Linear])
    pairs = [(r.get_idxs()[0], r.get_idxs()[1]) for r in rels]
    assert (id(m.c1), id(m.c2)) in pairs and (id(m.c2), id(m.c3)) in pairs


def test_module_called_twice_gets_two_nodes():
    class Twice(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(4, 4, 1); self.r = nn.ReLU()

        def forward(self, x):
            return self.r(self.c(self.r(x)))
    m = Twice()
    graph, bottoms = trace.trace_graph(m)
    relu_keys = [k for k, v in graph.items() if v is m.r]
    assert len(relu_keys) == 2 and relu_keys[0] == id(m.r) and isinstance(relu_keys[1], str)
    assert bottoms[id(m.c)] == [relu_keys[0]] and bottoms[relu_keys[1]] == [id(m.c)]


def test_traced_torchvision_mobilenetv2_calibrates_like_the_separate_calls(monkeypatch):
    """torchvision's MobileNetV2 (functional adaptive_avg_pool2d + flatten, `x + self.conv(x)` residuals, ReLU6) through the
    fx tracer and the one-residency plan == the drop-in calls one after the other (arithmetic by the oracle-backed fake)."""
    tv = pytest.importorskip("torchvision")
    import copy
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fakelib
    fakelib.install(monkeypatch)
    from dfq_b200 import dfq
    from dfq_b200.calibrate import GraphCalibration
    from dfq_b200.utils import layer_transform as LT
    from dfq_b200.utils.relation import create_relation
    torch.manual_seed(0)
    ma = tv.models.mobilenet_v2(width_mult=0.25).eval()
    for m in ma.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)

    def relu6_to_relu(parent):            # what the reference's --relu flag does (main_cls.py: switch_layers {ReLU6: ReLU});
        for name, ch in list(parent.named_children()):      # ReLU6 is not positively homogeneous, so create_relation stops at it
            if isinstance(ch, nn.ReLU6):
                setattr(parent, name, nn.ReLU())
            else:
                relu6_to_relu(ch)
    ga0, ba0 = trace.trace_graph(ma)
    assert len(create_relation(ga0, ba0, [nn.Conv2d, nn.Linear])) <= 1          # with ReLU6 in place nothing can be equalized
    relu6_to_relu(ma)
    mb = copy.deepcopy(ma)
    targ = [nn.Conv2d, nn.Linear]
    ga, ba = trace.trace_graph(ma)
    kinds = {type(v).__name__ if isinstance(v, nn.Module) else re.sub(r"_?\d+$", "", v) for v in ga.values()}
    assert {"Conv2d", "BatchNorm2d", "ReLU", "add", "AdaptiveAvgPool2d", "torch.flatten", "Dropout", "Linear"} <= kinds
    LT.merge_batchnorm(ma, ga, ba, targ)
    rels = create_relation(ga, ba, targ)
    assert len(rels) >= 30
    dfq.cross_layer_equalization(ga, rels, targ, converge_thres=1e-2)
    dfq.bias_correction(ga, ba, targ)
    gb, bb = trace.trace_graph(mb)
    cal = GraphCalibration(gb, bb, targ)
    res = cal.run(equalize=True, correction=True, converge_thres=1e-2)
    assert res.n_sweeps == dfq.cross_layer_equalization.last_result.n_sweeps and len(cal.relations) == len(rels)
    for (na, pa), (nb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert np.array_equal(pa.detach().numpy(), pb.detach().numpy()), na

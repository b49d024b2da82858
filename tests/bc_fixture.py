"""Inputs of the stand-alone bias-correction fixtures (tests/golden/ref_bc_<model>.npz).

Shared by tools/make_golden.py (which runs the REFERENCE's dfq.bias_correction on these inputs and stores what it
produced) and tests/test_gpu_entrypoints.py (which runs the CUDA path on the same inputs).  The inputs are regenerated
from the seed instead of being stored (MobileNetV2 + ResNet-18 weights are 60 MB): dfq_b200.workload.build_graph gives
seeded weights with imbalanced per-channel gains, and the BN buffers bias correction reads are registered directly
(fake_weight = |gamma|, fake_bias = beta, layer_transform.py:264-265) - no square root anywhere, so every machine
regenerates the same bits; the fixture carries the sha256 of every input tensor and the test refuses to run on a mismatch.
"""
import hashlib
import os

import numpy as np
import torch.nn as nn

from dfq_b200 import workload

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEEDS = {"resnet18": 13, "mobilenetv2": 17}
TARG = [nn.Conv2d, nn.Linear]


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def build(name):
    topo = workload.load_topology(os.path.join(GOLD, "topology_%s.json" % name))
    graph, bottoms, _ = workload.build_graph(topo, seed=SEEDS[name])
    for m in graph.values():
        if isinstance(m, nn.BatchNorm2d):
            m.register_buffer("fake_weight", m.weight.detach().abs().clone())
            m.register_buffer("fake_bias", m.bias.detach().clone())
    return graph, bottoms


def input_digests(graph):
    out = []
    for m in graph.values():
        if type(m) in TARG:
            out.append(sha(m.weight))
            if m.bias is not None:
                out.append(sha(m.bias))
        elif isinstance(m, nn.BatchNorm2d):
            out.append(sha(m.fake_weight)); out.append(sha(m.fake_bias))
    return np.array(out)

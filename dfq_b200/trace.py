"""graph / bottoms producer on torch.fx - SURVEY section 8(f) rank 3.

The calibration path consumes two dictionaries that the reference gets from its PyTransformer tracer (1.7-15 s per model on
the CPU, >99 % of the wall time once the arithmetic runs on the GPU; `TorchTransformer._build_graph`,
torchTransformer.py:485-602):

    graph[key]   = the nn.Module of a module node (key = id(module)), or the key string itself for a functional node
                   ("add_12", "iadd_31", "torch.cat_40", "torch.mean_77", "F.interpolate_5", "F.pad_9", "torch.flatten_80", ...)
    bottoms[key] = the keys of the nodes feeding it, in argument order (None for the root "Data")

`trace_graph(model)` builds the same dictionaries from a `torch.fx.symbolic_trace` of the model (tens of milliseconds), in
execution order, with the functional-op names the graph walks of this package key on (`'add' in key`, `'cat' in key`,
"F.pad", "torch.mean": utils/relation.py, graphwalk.py, utils/layer_transform.py).  Functional activations / pooling calls
are given module nodes (a fresh nn.ReLU / nn.AdaptiveAvgPool2d ...) because the walks classify nodes by module type.
A module instance called at several sites gets one node per call (key id(module) for the first, "<id>_<n>" after that).

Limits (same as symbolic tracing): data-dependent Python control flow in forward() is not traceable; pass
`concrete_args` for flags.  Nodes that do not carry an activation tensor (sizes, shapes, constants) are dropped.
"""
from __future__ import annotations

import operator
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

# functional call -> node-name stem used by the reference tracer's keys
_FUNC_NAMES = {
    operator.add: "add", torch.add: "add", operator.iadd: "iadd",
    torch.cat: "torch.cat", torch.mean: "torch.mean", torch.flatten: "torch.flatten",
    F.interpolate: "F.interpolate", F.pad: "F.pad", F.softmax: "F.softmax",
    operator.mul: "mul", torch.mul: "mul", torch.reshape: "torch.reshape",
}
_METHOD_NAMES = {"add": "add", "add_": "iadd", "mean": "torch.mean", "flatten": "torch.flatten", "view": "view",
                 "reshape": "torch.reshape", "contiguous": "contiguous", "mul": "mul", "permute": "permute", "squeeze": "squeeze"}
# functional forms of things the walks know as modules
_FUNC_AS_MODULE = {
    F.relu: lambda n: nn.ReLU(), torch.relu: lambda n: nn.ReLU(), F.relu6: lambda n: nn.ReLU6(),
    F.hardtanh: lambda n: nn.ReLU6() if (n.kwargs.get("min_val", n.args[1] if len(n.args) > 1 else -1.0) == 0.0
                                         and n.kwargs.get("max_val", n.args[2] if len(n.args) > 2 else 1.0) == 6.0) else nn.Hardtanh(),
    F.adaptive_avg_pool2d: lambda n: nn.AdaptiveAvgPool2d(n.args[1] if len(n.args) > 1 else n.kwargs["output_size"]),
    F.dropout: lambda n: nn.Dropout(), F.sigmoid: lambda n: nn.Sigmoid(), torch.sigmoid: lambda n: nn.Sigmoid(),
}
_NON_TENSOR_METHODS = {"size", "dim", "numel", "__len__"}
_NON_TENSOR_FUNCS = {getattr, operator.getitem, operator.floordiv, operator.mul, operator.sub, operator.truediv}


def trace_graph(model: nn.Module, concrete_args: Optional[dict] = None) -> Tuple["OrderedDict", "OrderedDict"]:
    """(graph, bottoms) of `model` in the reference tracer's format (see the module docstring)."""
    gm = fx.symbolic_trace(model, concrete_args=concrete_args)
    mods = dict(gm.named_modules())
    graph: "OrderedDict" = OrderedDict()
    bottoms: "OrderedDict" = OrderedDict()
    key_of: Dict[fx.Node, object] = {}        # fx node -> graph key, only for nodes that carry an activation
    calls: Dict[int, int] = {}
    counter = 0

    def inputs(node: fx.Node):
        seen = []

        def visit(a):
            if isinstance(a, fx.Node):
                if a in key_of:
                    seen.append(key_of[a])
            elif isinstance(a, (list, tuple)):
                for x in a:
                    visit(x)
            elif isinstance(a, dict):
                for x in a.values():
                    visit(x)
        visit(node.args)
        visit(node.kwargs)
        return seen

    def add(node: fx.Node, key, obj):
        key_of[node] = key
        graph[key] = obj
        bottoms[key] = inputs(node) or None

    n_inputs = 0
    for node in gm.graph.nodes:
        counter += 1
        if node.op == "placeholder":
            key = "Data" if n_inputs == 0 else "Data_%d" % n_inputs
            n_inputs += 1
            key_of[node] = key
            graph[key] = key
            bottoms[key] = None
        elif node.op == "call_module":
            m = mods[node.target]
            k = calls.get(id(m), 0)
            calls[id(m)] = k + 1
            add(node, id(m) if k == 0 else "%d_%d" % (id(m), k), m)
        elif node.op == "call_function":
            if node.target in _FUNC_AS_MODULE and inputs(node):
                m = _FUNC_AS_MODULE[node.target](node)
                add(node, id(m), m)
            elif node.target in _NON_TENSOR_FUNCS and not _carries_tensor(node, key_of):
                continue
            elif inputs(node):
                stem = _FUNC_NAMES.get(node.target, getattr(node.target, "__name__", "func"))
                key = "%s_%d" % (stem, counter)
                add(node, key, key)
        elif node.op == "call_method":
            if node.target in _NON_TENSOR_METHODS or not inputs(node):
                continue
            if node.target in ("relu", "relu_"):
                m = nn.ReLU()
                add(node, id(m), m)
                continue
            key = "%s_%d" % (_METHOD_NAMES.get(node.target, node.target), counter)
            add(node, key, key)
        # get_attr (parameters used functionally) and output carry nothing the walks need
    # keep the modules created for functional activations alive as long as the graph is
    graph_owner = [m for m in graph.values() if isinstance(m, nn.Module)]
    setattr(gm, "_dfq_graph_modules", graph_owner)
    trace_graph.last_module = gm
    return graph, bottoms


def _carries_tensor(node: fx.Node, key_of) -> bool:
    """operator.mul / getitem ... on activations are real nodes; on sizes they are bookkeeping."""
    if node.target in (getattr, operator.floordiv, operator.sub, operator.truediv):
        return False
    if node.target is operator.getitem:
        return False
    return all(isinstance(a, fx.Node) and a in key_of for a in node.args if isinstance(a, fx.Node)) and \
        any(isinstance(a, fx.Node) and a in key_of for a in node.args)

"""On-disk outputs adjacent to the calibration path.

  write_ncnn_table   the int8 calibration table the reference produces in convert_ncnn.py:178-201: one row per target
                     layer with the weight scale 128/max|W| repeated per output channel (`<name>_param_0 s s s ...`),
                     then one row per layer with the activation scale 128/max(|running_min|, |running_max|) (`<name> s`).
                     The per-tensor extrema are computed on the device (dfq_minmax).
  save_calibration   equalized / corrected state (weights, biases, fake BN statistics, scale vectors) as one .pt file.
"""
import torch

from .utils.quantize import tensor_minmax


def ncnn_scales(graph, targ_type):
    """[(weight_scale, out_channels, activation_scale | None)] per target layer, in graph order."""
    rows = []
    for key in graph:
        layer = graph[key]
        if type(layer) not in targ_type:
            continue
        mm = tensor_minmax(layer.weight.detach()).tolist()                     # convert_ncnn.py:186-187
        w_scale = 128. / max(abs(mm[1]), abs(mm[0]))
        a_scale = None
        if hasattr(layer, "quant"):
            mi, ma = float(torch.min(layer.quant.running_min)), float(torch.max(layer.quant.running_max))
            a_scale = 128. / max(abs(ma), abs(mi))                            # :189-191
        rows.append((w_scale, layer.weight.shape[0], a_scale))
    return rows


def write_ncnn_table(graph, path, targ_type, names=None):
    """Write the table; `names` are the ncnn blob names of the target layers (default: layer_<i>)."""
    rows = ncnn_scales(graph, targ_type)
    names = names or ["layer_%d" % i for i in range(len(rows))]
    with open(path, "w") as f:
        for n, (ws, oc, _) in zip(names, rows):
            f.write(' '.join(["%s_param_0" % n] + [str(ws)] * oc) + '\n')
        for n, (_, _, a) in zip(names, rows):
            if a is not None:
                f.write("%s %s\n" % (n, str(a)))
    return rows


def save_calibration(graph, relations, path):
    """Persist what the passes changed: parameters of every module with weights, fake BN statistics and Relation.S."""
    state = {"layers": {}, "bn": {}, "S": [None if r.S is None else r.S.detach().cpu() for r in relations]}
    for i, key in enumerate(graph):
        m = graph[key]
        if isinstance(m, str):
            continue
        if hasattr(m, "fake_weight"):
            state["bn"][i] = {"fake_weight": m.fake_weight.detach().cpu(), "fake_bias": m.fake_bias.detach().cpu()}
        elif hasattr(m, "weight") and m.weight is not None:
            state["layers"][i] = {"weight": m.weight.detach().cpu(), "bias": None if m.bias is None else m.bias.detach().cpu()}
    torch.save(state, path)
    return state

"""Graph surgery around the calibration path: drop-in for the reference's ``utils/layer_transform.py``.

Public names, signatures, module-level state and prints follow jakc4103/DFQ:

    replace_op / restore_op, patched ops, CustomTensorOP, switch_layers   layer_transform.py:16-228
    merge_batchnorm            layer_transform.py:231-276  -> dfq_bn_fold          (device)
    quantize_targ_layer        layer_transform.py:279-296  -> dfq_quantize_tensors (device)
    find_prev_bn               layer_transform.py:299-344  (graph walk, host)
    set_quant_minmax           layer_transform.py:347-609  (per-channel [C] vectors, host: SURVEY.md a12)

The op patching keeps the reference's protocol: a functional op is quantized when it is called from a
``forward`` whose line number matches the next recorded op name (``add_<line>_2``, ``torch_cat_<line>_<n>``,
...), which is how observers are attached to the UNMODIFIED model files.
"""
import sys
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..engine import Session
from ..graphwalk import find_prev_bn, merge_order  # noqa: F401  (find_prev_bn is part of this module's API)
from .quantize import QConv2d, QuantConv2d, QuantNConv2d, QLinear, QuantLinear, QuantNLinear, quantize, QuantMeasure  # noqa: F401

tensor_target = torch.Tensor
raw_tensor_magic_op = {}
tensor_magic_op_supported = ['__add__', 'add', '__iadd__']
raw_torch_op = {}
torch_op_supported = ['cat', 'mean']
raw_func_op = {}
func_op_sopprted = ['interpolate', 'softmax']
module_tensor_op = None


def _called_from_forward_as(template, *fmt):
    """True when the patched op was called directly from a `forward` and the recorded name of the next
    functional op equals template.format(<caller line>, ...)  (layer_transform.py:19-20,58-59,...)."""
    caller = sys._getframe(2)
    if caller.f_code.co_name != 'forward' or module_tensor_op is None:
        return False
    return template.format(caller.f_lineno, *fmt) == module_tensor_op.get_module_name()


def ___add__(input, *args):
    if _called_from_forward_as('add_{}_2'):
        input = module_tensor_op(input)
        args = [module_tensor_op(args[0])]
        module_tensor_op.add_idx_name_tensor_op()
    return raw_tensor_magic_op['__add__'](input, *args)


def ___iadd__(input, *args):
    # quirk Q5 (layer_transform.py:44): the in-place add is executed as an out-of-place __add__
    if _called_from_forward_as('iadd_{}_2'):
        input = module_tensor_op(input)
        args = [module_tensor_op(args[0])]
        module_tensor_op.add_idx_name_tensor_op()
    return raw_tensor_magic_op['__add__'](input, *args)


def _add(input, *args):
    # Tensor.add goes through ___add__ one frame deeper, so (as in the reference) it never matches `forward`
    return ___add__(input, *args)


def torch_cat(inputs, dim=0):
    if _called_from_forward_as('torch_cat_{}_{}', len(inputs)):
        inputs = [module_tensor_op(t) for t in inputs]
        module_tensor_op.add_idx_name_tensor_op()
    return raw_torch_op['cat'](tuple(inputs), dim)


def torch_mean(input, dim=None, keepdim=False, out=None):
    if _called_from_forward_as('torch_mean_{}_1'):
        input = module_tensor_op(input)
        module_tensor_op.add_idx_name_tensor_op()
    if dim is None:
        return raw_torch_op['mean'](input)
    return raw_torch_op['mean'](input, dim)


def F_interpolate(input, size=None, scale_factor=None, mode='nearest', align_corners=None):
    if _called_from_forward_as('F_interpolate_{}_1'):
        input = module_tensor_op(input)
        module_tensor_op.add_idx_name_tensor_op()
    return raw_func_op['interpolate'](input, size, scale_factor, mode, align_corners)


def F_softmax(input, dim=None, _stacklevel=3, dtype=None):
    if _called_from_forward_as('F_softmax_{}_1'):
        input = module_tensor_op(input)
        module_tensor_op.add_idx_name_tensor_op()
    return raw_func_op['softmax'](input, dim, _stacklevel, dtype)


def replace_op():
    for op_name in tensor_magic_op_supported:
        raw_tensor_magic_op[op_name] = getattr(torch.Tensor, op_name)
        setattr(tensor_target, op_name, globals()['_' + op_name])
    for op_name in torch_op_supported:
        raw_torch_op[op_name] = getattr(torch, op_name)
        setattr(torch, op_name, globals()['torch_' + op_name])
    for op_name in func_op_sopprted:
        raw_func_op[op_name] = getattr(F, op_name)
        setattr(F, op_name, globals()['F_' + op_name])


def restore_op():
    for op_name in tensor_magic_op_supported:
        setattr(tensor_target, op_name, raw_tensor_magic_op[op_name])
    for op_name in torch_op_supported:
        setattr(torch, op_name, raw_torch_op[op_name])
    for op_name in func_op_sopprted:
        setattr(F, op_name, raw_func_op[op_name])


def switch_layers(model, transformer, data, module_dict, ignore_layer=[], ignore_op=['pad'], quant_op=True):
    # swap layer classes, then trace (the graph must be built after every state_dict is in place)
    for key in module_dict:
        for source, target in module_dict[key]:
            transformer.register(source, target)
        model = transformer.trans_layers(model, update=True if key == 1 else False)
    transformer._build_graph(model, data, ignore_layer)
    if not quant_op:
        return model, transformer

    global module_tensor_op
    recorded = [rec for rec in transformer.log.getRecordTensorOP() if not any(ig in rec[1] for ig in ignore_op)]
    observers = []
    for _, op_name in recorded:
        observers.extend(QuantMeasure(num_bits=8, momentum=0.1) for _ in range(int(op_name.split('_')[-1])))
    module_tensor_op = CustomTensorOP(observers, recorded)
    model.add_module('custom_tensor_op', module_tensor_op)
    setattr(model, 'name_tensor_op', recorded)
    setattr(model, 'idx_name_tensor_op', 0)
    setattr(model, 'idx_tensor_op', 0)
    return model, transformer


class CustomTensorOP(nn.Module):
    """
    special module used for quantization of torch.xxx(), F.xxx() and torch.Tensor.__xxx__()
    """

    def __init__(self, tensor_op, name_tensor_op):
        super(CustomTensorOP, self).__init__()
        for idx, op in enumerate(tensor_op):
            self.add_module(str(idx), op)
        self.idx_tensor_op = 0
        self.len = len(tensor_op)
        self.name_tensor_op = name_tensor_op
        self.idx_name_tensor_op = 0
        self.num_op = len(name_tensor_op)

    def add_idx_tensor_op(self):
        self.idx_tensor_op = (self.idx_tensor_op + 1) % self.len

    def add_idx_name_tensor_op(self):
        self.idx_name_tensor_op = (self.idx_name_tensor_op + 1) % self.num_op

    def get_module_name(self):
        return self.name_tensor_op[self.idx_name_tensor_op][1]

    def get_graph_name(self):
        return self.name_tensor_op[self.idx_name_tensor_op][0]

    def get_module_next(self):
        mod = self._modules[str(self.idx_tensor_op)]
        self.add_idx_tensor_op()
        return mod

    def forward(self, x):
        x = self._modules[str(self.idx_tensor_op)](x)
        self.add_idx_tensor_op()
        return x


def _identity_bn_eps():
    """layer_transform.py:272 sets eps = 0; torch >= 2.x rejects that in F.batch_norm, and 1e-12 gives the
    bit-identical identity (1/sqrt(1 + 1e-12) == 1.0f)."""
    try:
        F.batch_norm(torch.zeros(1, 1), torch.zeros(1), torch.ones(1), None, None, False, 0.0, 0.0)
        return 0
    except Exception:
        return 1e-12


def merge_batchnorm(model, graph, bottoms, targ_type=[QConv2d]):
    """!
    This function will merge params and stats of BatchNorm into targ_type like QuantConv2d.
    Once the values is merged, the values of layer will be set to default (as an identity layer),
    and it creates buffer named 'fake_weight' adn 'fake_bias' for latter usage of set_quant_minmax
    """
    with torch.no_grad():
        pairs = []
        for key in graph:
            if bottoms[key] is None:
                continue
            for src in bottoms[key]:
                if type(graph[key]) == nn.BatchNorm2d and type(graph[src]) in targ_type:
                    pairs.append((key, src))
                    break                                   # only the first matching input is folded (:274)
        if not pairs:
            return model
        sess = Session()
        folds = []
        for bn_key, conv_key in pairs:
            bn, conv = graph[bn_key], graph[conv_key]
            if conv.bias is None:                           # :253-254
                conv.bias = nn.Parameter(data=torch.zeros((conv.weight.size(0)), dtype=torch.float32,
                                                          device=conv.weight.device), requires_grad=False)
            li = sess.add_layer(conv.weight, conv.bias)
            n = conv.weight.size(0)
            folds.append(dict(layer=li, bn_eps=bn.eps,
                              gamma_off=sess.bind(bn.weight.detach(), False), beta_off=sess.bind(bn.bias.detach(), False),
                              mean_off=sess.bind(bn.running_mean, False), var_off=sess.bind(bn.running_var, False),
                              fake_w_off=sess.alloc(n), fake_b_off=sess.alloc(n)))
        sess.upload()
        sess.run_bn_fold(folds)
        fakes = [(sess.view(f["fake_w_off"], sess.layer(f["layer"])["rows"]).clone(),
                  sess.view(f["fake_b_off"], sess.layer(f["layer"])["rows"]).clone()) for f in folds]
        sess.download()
        eps = _identity_bn_eps()
        for (bn_key, _), (fw, fb) in zip(pairs, fakes):
            bn = graph[bn_key]
            # store values for later usage. ex: set_quant_min_max and bias correction (:264-265)
            bn.register_buffer('fake_weight', fw.to(bn.weight.device))
            bn.register_buffer('fake_bias', fb.to(bn.weight.device))
            # the batch norm becomes an identity layer (:268-272)
            bn.weight.fill_(1)
            bn.running_var.fill_(1)
            bn.bias.fill_(0)
            bn.running_mean.fill_(0)
            bn.eps = eps
    return model


def quantize_targ_layer(graph, bit_weight=8, bits_bias=16, targ_type=None):
    print("Quantizing Layer parameters")
    if bits_bias == 32:
        print("Skipping bias quantization (32 bits)")
    assert targ_type != None, "targ_type cannot be None!"
    with torch.no_grad():
        sess = Session()
        tasks = []
        on_cuda = None
        for key in graph:
            layer = graph[key]
            if type(layer) not in targ_type:
                continue
            if on_cuda is None:
                on_cuda = layer.weight.is_cuda
            tasks.append((sess.bind(layer.weight), layer.weight.numel(), bit_weight, False))
            if layer.bias is not None and bits_bias < 32:
                tasks.append((sess.bind(layer.bias), layer.bias.numel(), bits_bias, False))
        if tasks:
            sess.upload()
            # "quantization behave differently on cpu and gpu" (:287): CPU parameters get the true division
            sess.run_quantize(tasks, div_mode=1 if on_cuda else 0)
            sess.download()
    return graph


def set_quant_minmax(graph, bottoms, is_detection=False, bn_type=torch.nn.BatchNorm2d, N=6, verbose=True):
    """!
    This function set the running_min/running_max value of QuantMeasure using the statistics form previous BatchNorm layer.

    Cases (layer_transform.py:348-370): (a) one BN per observer, (b) one observer fed through an add/cat of several
    BNs, (c) several observers each fed by several BNs, (d) a conv/linear without BN in between (SSD heads).
    Element-wise additions accumulate means and variances of (rectified) Gaussians; concatenations take the
    min/max over branches; otherwise the min/max of the inputs are averaged.  Works on [C] vectors: host code.
    """
    from scipy.stats import norm
    if verbose:
        print("SET QUANT MIN MAX")

    def observers_of(layer):
        if type(layer) == str:
            if module_tensor_op.get_graph_name() == layer:
                tags = [n.replace('_', '') for n in tensor_magic_op_supported + torch_op_supported + func_op_sopprted]
                if any(tag in layer for tag in tags):
                    count = int(module_tensor_op.get_module_name().split('_')[-1])
                    mods = [module_tensor_op.get_module_next() for _ in range(count)]
                    module_tensor_op.add_idx_name_tensor_op()
                    return mods
        elif hasattr(layer, 'quant'):
            return [getattr(layer, 'quant')]
        return None

    eps = 1e-6
    hi = lambda b, w, n: float(torch.max(b + n * w))
    lo = lambda b, w, n: float(torch.min(b - n * w))
    pdf = lambda x: torch.from_numpy(norm(0, 1).pdf(x)).float()
    cdf = lambda x: torch.from_numpy(norm.cdf(x)).float()
    # moments of max(0, X) and of min(6, max(0, X)) for X ~ N(bias, weight^2)  (:411-422)
    mean_relu = lambda w, b: w * pdf(-b / w) + b * (1 - cdf(-b / w))
    var_relu = lambda w, b, m: (1 - cdf(-b / w)) * (b * b + w * w + m * m - 2 * m * b) + \
        w * (b - 2 * m) * (pdf(-b / w)) + \
        m * m * cdf(-b / w)
    mean_relu6 = lambda w, b: w * (pdf(-b / w) - pdf((6 - b) / w)) + \
        b * (cdf((6 - b) / w) - cdf(-b / w)) + \
        6 * (1 - cdf((6 - b) / w))
    var_relu6 = lambda w, b, m: (cdf((6 - b) / w) - cdf(-b / w)) * (b * b + w * w + m * m - 2 * m * b) + \
        w * (-6) * pdf((6 - b) / w) + \
        w * (b - 2 * m) * (pdf(-b / w) - pdf((6 - b) / w)) + \
        m * m * cdf(-b / w) + \
        ((6 - m) ** 2) * (1 - cdf((6 - b) / w))

    def clipped_range(b, w, act):
        vmin = max(0., lo(b, w, N)) if 'relu' in act else lo(b, w, N)
        vmax = min(6., hi(b, w, N)) if 'relu6' in act else hi(b, w, N)
        return vmin, vmax

    def moments(w, b, act):
        if 'relu' == act:
            m = mean_relu(w, b)
            return m, var_relu(w, b, m)
        if 'relu6' == act:
            m = mean_relu6(w, b)
            return m, var_relu6(w, b, m)
        return b, w * w

    bn_module, relu_attached = {}, {}
    for key in graph:
        bot = bottoms[key]
        if bot is None:
            continue
        node = graph[key]
        if type(node) == bn_type:
            bn_module[key] = node
            relu_attached[key] = 'none'
            continue
        if type(node) == torch.nn.ReLU:
            relu_attached[bot[0]] = 'relu'
        elif type(node) == torch.nn.ReLU6:
            relu_attached[bot[0]] = 'relu6'

        quant_module = observers_of(node)
        if len(bot) == 1 and bot[0] == 'Data':
            if is_detection:
                quant_module[0].running_max.fill_(1)
                quant_module[0].running_min.fill_(-1)
            else:  # (1 - mean)/std and (0 - mean)/std of the ImageNet preprocessing (:448-449)
                quant_module[0].running_max.fill_(2.64)
                quant_module[0].running_min.fill_(-2.11790393)
            continue
        if quant_module is None:
            continue

        bn_list, act_list, type_list, targ_without_bn = find_prev_bn(bn_module, relu_attached, graph, bottoms, bot[:])
        if len(quant_module) == len(bn_list):                       # case (a): 1 to 1
            for qm, (bn, bid), act in zip(quant_module, bn_list, act_list):
                bias = getattr(bn, 'fake_bias').view(-1)
                weight = getattr(bn, 'fake_weight').view(-1)
                if bid[0] in targ_without_bn:                       # case (d) (:459-475)
                    layer_type, obj = targ_without_bn[bid[0]]
                    lw = getattr(obj, 'weight').detach().data
                    lb = getattr(obj, 'bias').detach().data
                    if layer_type == 'conv':
                        lw = lw.view(lw.size(0), lw.size(1), -1).sum(-1)
                        lw = lw.view(lw.size(0), lw.size(1), 1, 1)
                        groups = getattr(obj, 'groups')
                        bias = F.conv2d(bias.view(1, -1, 1, 1), lw, lb, 1, 0, 1, groups)
                        weight = F.conv2d(weight.view(1, -1, 1, 1), lw, lb, 1, 0, 1, groups)
                    else:
                        bias = F.linear(bias.view(1, -1), lw, lb)
                        weight = F.linear(weight.view(1, -1), lw, lb)
                    vmax, vmin = hi(bias, weight, N), lo(bias, weight, N)
                else:
                    vmin, vmax = clipped_range(bias, weight, act)
                qm.running_max.fill_(vmax)
                qm.running_min.fill_(vmin)
            continue

        # cases (b)/(c): fold the BNs of each direct input
        branches = OrderedDict()
        for (bn, bid), act, ctype in zip(bn_list, act_list, type_list):
            branches.setdefault(bid[0], []).append((bid, bn, act, ctype))
        results = {}
        for bkey, entries in branches.items():
            ordered = merge_order(entries)
            bid0, bn0, act0, ctype = ordered[0]
            depth = len(bid0)
            rest = ordered[1:]
            bias = bn0.fake_bias.detach().clone()
            weight = bn0.fake_weight.detach().clone()
            if 'add' in ctype:
                mean, var = moments(weight, bias, act0)
            else:
                vmin, vmax = clipped_range(bias, weight, act0)
            while rest:
                run = 0
                while run < len(rest) and len(rest[run][0]) == depth:
                    run += 1
                if run == 0:
                    depth = len(rest[0][0])               # nothing left at this depth: cut
                    continue
                for _, bn_t, act_t, ctype in rest[:run]:
                    bias = bn_t.fake_bias.detach().clone()
                    weight = bn_t.fake_weight.detach().clone()
                    if 'add' in ctype:
                        m_t, v_t = moments(weight, bias, act_t)
                        if act_t in ('relu', 'relu6'):
                            mean += m_t
                            var += v_t
                        else:
                            mean += bias
                            var += weight * weight
                        # an activation after the add re-rectifies the running sum (:545-552)
                        if 'relu6' in ctype:
                            prev = mean
                            mean = mean_relu6(torch.sqrt(var + eps), mean)
                            var = var_relu6(torch.sqrt(var + eps), prev, mean)
                        elif 'relu' in ctype:
                            prev = mean
                            mean = mean_relu(torch.sqrt(var + eps), mean)
                            var = var_relu(torch.sqrt(var + eps), prev, mean)
                    elif 'cat' == ctype:
                        vmin = min(vmin, max(0., lo(bias, weight, N)) if 'relu' in act_t else lo(bias, weight, N))
                        vmax = max(vmax, min(6., hi(bias, weight, N)) if 'relu6' in act_t else hi(bias, weight, N))
                    else:
                        vmin += max(0., lo(bias, weight, N)) if act_t else lo(bias, weight, N)
                        vmax += hi(bias, weight, N)
                rest = rest[run:]
                if 'one' == ctype:
                    vmin /= (run + 1)
                    vmax /= (run + 1)
            results[bkey] = (ctype, mean, var) if 'add' in ctype else (ctype, vmin, vmax)

        def final_range(res):
            if 'add' in res[0]:
                _, mean, var = res
                return lo(mean, torch.sqrt(var + eps), N), hi(mean, torch.sqrt(var + eps), N)
            return res[1], res[2]

        if len(quant_module) == 1 and len(quant_module) < len(bn_list):          # (b) 1 to many
            assert len(list(results.keys())) == 1, "Error occurs when setting min/max, should be 1 to many"
            vmin, vmax = final_range(list(results.values())[0])
            quant_module[0].running_max.fill_(vmax)
            quant_module[0].running_min.fill_(vmin)
        elif len(quant_module) < len(bn_list):                                   # (c) many to many
            assert len(results) == len(quant_module), 'LENGTH NOT EQUAL {} vs {}'.format(len(results), len(quant_module))
            for idx in range(len(results)):
                vmin, vmax = final_range(results[str(idx)])
                quant_module[idx].running_max.fill_(vmax)
                quant_module[idx].running_min.fill_(vmin)
        else:
            assert False, "Unknown error occured while setting min/max"

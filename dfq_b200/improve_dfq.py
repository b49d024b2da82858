"""Distilled-data range update: drop-in for the two live functions of the reference's ``improve_dfq.py``.

    update_quant_range   improve_dfq.py:280-297   forward the distilled batches with every QuantMeasure in
                                                   update_stat mode (observers run on the GPU, no host syncs), then pin the
                                                   first layer's range to the preprocessing constants
    set_update_stat      improve_dfq.py:299-309   toggle update_stat on every module of the given types

    GradHook, ModuleHook improve_dfq.py:12-100    hook holders; `ZeroQ/distill_data.py:28` imports GradHook by name

The remaining names of improve_dfq.py (update_scale, set_scale, transform_quant_layer, bias_correction_distill) are the
author's abandoned learned-scale experiment (README.md:194-195); the main scripts import them but every call site is
commented out.  The names exist (the unmodified scripts import them) and raise NotImplementedError when called - they are
outside the calibration path.
"""
import torch

from .utils.layer_transform import replace_op, restore_op


def update_quant_range(model, data, graph, bottoms, is_detection=False):
    with torch.no_grad():
        replace_op()
        try:
            for batch in data:
                _ = model(batch.cuda())
        finally:
            restore_op()
        for key in graph:
            if bottoms[key] is None:
                continue
            if bottoms[key][0] == "Data":
                quant = graph[key].quant
                if not is_detection:
                    quant.running_max.fill_(2.64)
                    quant.running_min.fill_(-2.11790393)
                else:
                    quant.running_max.fill_(1)
                    quant.running_min.fill_(-1)
    return model


def set_update_stat(model, targ_type, update_stat):
    """!
    this function turns on/off the update_stat flag in modules in targ_type
    """
    for name, child in model._modules.items():
        if len(child._modules) > 0 and type(child) not in targ_type:
            set_update_stat(child, targ_type, update_stat)
        elif type(child) in targ_type:
            child.set_update_stat(update_stat)


class ModuleHook(object):
    """Forward hook that remembers the module it fired on together with that call's inputs and outputs
    (improve_dfq.py:82-100; imported by name by the reference's experiments)."""

    def __init__(self):
        self.module = self.inputs = self.outputs = None

    def hook(self, module, input, output):
        self.module, self.inputs, self.outputs = module, input, output

    def clear(self):
        self.module = self.inputs = self.outputs = None


class GradHook(object):
    """Holder of a weight and its optional (prev-)scale vectors with pass-through gradient hooks
    (improve_dfq.py:12-80).  `ZeroQ/distill_data.py:28` imports the name; every use of it there is commented out and both
    gradient hooks of the reference return their argument unchanged on their first line - that live behaviour is what is
    kept; the outlier mask is still computed so `.mask` exists."""

    def __init__(self, weight, scale=None, scale_prev=None, merge_scale=None, merge_scale_prev=None):
        self.weight, self.scale, self.scale_prev = weight, scale, scale_prev
        self.merge_scale, self.merge_scale_prev = merge_scale, merge_scale_prev
        self.update_mask()

    def get_weight_scaled(self):
        w = self.weight
        if self.scale_prev is not None:
            w = self.merge_scale_prev(w, self.scale_prev)
        if self.scale is not None:
            w, _ = self.merge_scale(w, None, self.scale)
        return w

    def update_mask(self):
        with torch.no_grad():
            # clone: get_weight_scaled returns the parameter itself when no scale is attached, and the masking is in place
            w = self.get_weight_scaled().detach().clone()
            mean, std = w.mean(), torch.sqrt(torch.var(w))
            # improve_dfq.py:34-35 indexes with an integer tensor (values 1 or 2: inside one or both 2-sigma bounds), i.e.
            # it zeroes the slices w[1] and w[2] whenever they are addressed - reproduced literally
            idx = (w < (mean + 2 * std)).long() + (w > (mean - 2 * std)).long()
            w[idx] = 0
            self.mask = torch.abs(w) / torch.abs(w).max()

    def hook_mask_grad_tensor(self, grad):
        return grad

    def hook_mask_grad_input(self, m, grad_input, grad_output):
        return grad_input


def _not_on_the_path(name):
    def call(*args, **kwargs):
        raise NotImplementedError(
            "%s is part of the reference's abandoned learned-scale experiment (README.md:194-195, every call site in "
            "main_cls/main_seg/main_ssd is commented out) and is outside the calibration path dfq_b200 implements "
            "(SURVEY.md section 2, row 6); the name exists so that the unmodified main scripts import." % name)
    call.__name__ = name
    return call


update_scale = _not_on_the_path("update_scale")
set_scale = _not_on_the_path("set_scale")
transform_quant_layer = _not_on_the_path("transform_quant_layer")
bias_correction_distill = _not_on_the_path("bias_correction_distill")

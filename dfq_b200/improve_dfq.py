"""Distilled-data range update: drop-in for the two live functions of the reference's ``improve_dfq.py``.

    update_quant_range   improve_dfq.py:280-297   forward the distilled batches with every QuantMeasure in
                                                   update_stat mode (observers run on the GPU, no host syncs), then pin the
                                                   first layer's range to the preprocessing constants
    set_update_stat      improve_dfq.py:299-309   toggle update_stat on every module of the given types

The remaining names of improve_dfq.py (update_scale, set_scale, transform_quant_layer, bias_correction_distill) are the
author's abandoned learned-scale experiment (README.md:194-195); the main scripts import them but never call them on the
supported flag combinations.  They are forwarded to the reference implementation when it is importable and raise
otherwise - they are outside the calibration path.
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


def _forward_to_reference(name):
    def call(*args, **kwargs):
        import importlib.util
        import os
        root = os.environ.get("DFQ_REFERENCE_ROOT", "/root/reference")
        path = os.path.join(root, "improve_dfq.py")
        if not os.path.isfile(path):
            raise NotImplementedError("%s belongs to the reference's abandoned learned-scale experiment and is not part "
                                      "of the calibration path; reference tree not found at %s" % (name, root))
        spec = importlib.util.spec_from_file_location("_ref_improve_dfq", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return getattr(mod, name)(*args, **kwargs)
    call.__name__ = name
    return call


update_scale = _forward_to_reference("update_scale")
set_scale = _forward_to_reference("set_scale")
transform_quant_layer = _forward_to_reference("transform_quant_layer")
bias_correction_distill = _forward_to_reference("bias_correction_distill")

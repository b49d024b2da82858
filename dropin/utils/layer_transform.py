"""`import utils.layer_transform` shim."""
from dfq_b200.utils.layer_transform import (switch_layers, replace_op, restore_op, set_quant_minmax, merge_batchnorm,  # noqa: F401
                                            quantize_targ_layer, find_prev_bn, CustomTensorOP)

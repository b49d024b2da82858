"""`import dfq` shim: puts dfq_b200's implementation behind the reference's module path."""
from dfq_b200.dfq import (cross_layer_equalization, bias_absorption, bias_correction, _quantize_error, clip_weight,  # noqa: F401
                          _layer_equalization)

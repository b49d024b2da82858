"""`import utils.quantize` shim."""
from dfq_b200.utils.quantize import (UniformQuantize, quantize, QuantMeasure, QConv2d, QuantConv2d, QuantNConv2d,  # noqa: F401
                                     QLinear, QuantLinear, QuantNLinear, set_layer_bits)

"""`import improve_dfq` shim."""
from dfq_b200.improve_dfq import (update_scale, transform_quant_layer, set_scale, update_quant_range, set_update_stat,  # noqa: F401
                                  bias_correction_distill, GradHook, ModuleHook)

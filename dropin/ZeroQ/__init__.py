"""`import ZeroQ` shim: `ZeroQ.distill_data.getDistilData` comes from dfq_b200; the rest of the package
(ZeroQ.utils: random-data loaders, quantization helpers - outside the path) still resolves to the reference tree."""
import os

_ref = os.path.join(os.environ.get("DFQ_REFERENCE_ROOT", "/root/reference"), "ZeroQ")
if os.path.isdir(_ref):
    __path__.append(_ref)

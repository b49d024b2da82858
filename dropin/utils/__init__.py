"""`import utils` shim.  The calibration modules come from dfq_b200; everything else the main scripts import from
`utils` (metrics, segmentation, detection helpers - evaluation code, outside the path) still resolves to the
reference tree when it is present."""
import os

from dfq_b200.utils import visualize_per_layer  # noqa: F401

_ref_utils = os.path.join(os.environ.get("DFQ_REFERENCE_ROOT", "/root/reference"), "utils")
if os.path.isdir(_ref_utils):
    __path__.append(_ref_utils)

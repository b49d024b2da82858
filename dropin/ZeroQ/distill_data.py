"""`from ZeroQ.distill_data import getDistilData` shim."""
from dfq_b200.distill import getDistilData  # noqa: F401

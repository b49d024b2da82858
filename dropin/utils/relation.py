"""`import utils.relation` shim."""
from dfq_b200.utils.relation import Relation, create_relation  # noqa: F401

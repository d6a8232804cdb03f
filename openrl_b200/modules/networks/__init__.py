from .policy_network import PolicyNetwork  # noqa: F401
from .value_network import ValueNetwork  # noqa: F401

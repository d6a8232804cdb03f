from .policy_network import PolicyNetwork  # noqa: F401
from .value_network import ValueNetwork  # noqa: F401
from .policy_value_network import PolicyValueNetwork  # noqa: F401

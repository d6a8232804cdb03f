from .onpolicy_driver import OnPolicyDriver  # noqa: F401

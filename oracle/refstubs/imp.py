"""py3.12 shim for the removed `imp` module (reference: openrl/envs/mpe/scenarios/__init__.py:1,7)."""
import importlib.util


def load_source(name, pathname):
    spec = importlib.util.spec_from_file_location(name, pathname)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

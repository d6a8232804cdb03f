"""argparse-backed stand-in for jsonargparse (test infrastructure only; lets the
reference's openrl/configs/config.py:24 parser be built unmodified)."""
import argparse
import typing

import yaml


def _str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("true", "1", "yes", "y", "t"):
        return True
    if v.lower() in ("false", "0", "no", "n", "f"):
        return False
    raise argparse.ArgumentTypeError(f"bool expected, got {v!r}")


def _yaml_value(v):
    return yaml.safe_load(v) if isinstance(v, str) else v


class Namespace(argparse.Namespace):
    def __getitem__(self, k):
        return getattr(self, k)

    def __contains__(self, k):
        return hasattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)


def _nest(ns):
    """dotted keys -> nested namespaces (reward_class.id -> cfg.reward_class.id)."""
    out = Namespace()
    for k, v in vars(ns).items():
        if "." in k:
            head, tail = k.split(".", 1)
            sub = getattr(out, head, None)
            if not isinstance(sub, Namespace):
                sub = Namespace()
                setattr(out, head, sub)
            setattr(sub, tail, v)
        else:
            setattr(out, k, v)
    return out


class ActionConfigFile(argparse.Action):
    def __init__(self, option_strings, dest, **kwargs):
        kwargs.pop("type", None)
        super().__init__(option_strings, dest, **kwargs)

    def __call__(self, parser, cfg, values, option_string=None):
        with open(values) as f:
            data = yaml.safe_load(f) or {}

        def put(prefix, d):
            for k, v in d.items():
                key = f"{prefix}{k}"
                if isinstance(v, dict) and any(a.dest.startswith(key + ".") for a in parser._actions):
                    put(key + ".", v)
                else:
                    setattr(cfg, key, v)

        put("", data)


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *args, **kwargs):
        kwargs.pop("env_prefix", None)
        kwargs.pop("default_env", None)
        super().__init__(*args, **kwargs)

    def add_argument(self, *names, **kwargs):
        t = kwargs.get("type")
        if t is bool:
            kwargs["type"] = _str2bool
        elif t is dict or typing.get_origin(t) in (list, typing.List) or t in (list,):
            kwargs["type"] = _yaml_value
        # reference config.py:1045-1065 declares four dash-less option names
        if len(names) == 1 and not names[0].startswith("-") and ("default" in kwargs or "type" in kwargs):
            names = ("--" + names[0],)
        return super().add_argument(*names, **kwargs)

    def parse_args(self, args=None, namespace=None):
        ns = super().parse_args(args, namespace if namespace is not None else Namespace())
        return _nest(ns)

"""Minimal stand-in for gin-config (TEST INFRASTRUCTURE, oracle/ only).

Supports exactly the syntax the reference's seven .gin files use: ``Class.field = <python literal>``
and ``#`` comments."""
import ast

_BINDINGS = {}


def clear_config():
    _BINDINGS.clear()


def parse_config_file(path):
    with open(path) as fh:
        for raw in fh:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            lhs, rhs = line.split("=", 1)
            cls, field = lhs.strip().split(".", 1)
            _BINDINGS.setdefault(cls, {})[field.strip()] = ast.literal_eval(rhs.strip())


def configurable(cls):
    name = cls.__name__
    orig_init = cls.__init__

    def __init__(self, *args, **kwargs):
        merged = dict(_BINDINGS.get(name, {}))
        merged.update(kwargs)
        orig_init(self, *args, **merged)

    cls.__init__ = __init__
    return cls

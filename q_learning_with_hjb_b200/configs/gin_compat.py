"""``gin`` if gin-config is installed, otherwise a small built-in stand-in.

The reference binds its config dataclasses with gin-config (``@gin.configurable`` +
``gin.parse_config_file``; scripts/test_vhjb_policy.py:51-55).  The only syntax its seven .gin files use is
``Class.field = <python literal>`` plus ``#`` comments, which the stand-in below handles, so the same
files load whether or not gin-config is available.
"""
from __future__ import annotations

import ast
import functools

try:  # pragma: no cover - exercised only where gin-config is installed
    import gin as _real_gin
    _HAVE_GIN = hasattr(_real_gin, "parse_config_file") and hasattr(_real_gin, "configurable") \
        and "oracle/_stubs" not in (getattr(_real_gin, "__file__", "") or "")
except ImportError:
    _real_gin = None
    _HAVE_GIN = False

if _HAVE_GIN:
    configurable = _real_gin.configurable
    parse_config_file = _real_gin.parse_config_file
    clear_config = _real_gin.clear_config
else:
    _bindings: dict = {}

    def clear_config():
        _bindings.clear()

    def parse_config(text: str):
        for lineno, raw in enumerate(text.splitlines(), 1):
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if "=" not in line or "." not in line.split("=", 1)[0]:
                raise ValueError(f"gin_compat: cannot parse line {lineno}: {raw!r}")
            target, value = line.split("=", 1)
            scope, field = target.strip().rsplit(".", 1)
            _bindings.setdefault(scope.split("/")[-1], {})[field] = ast.literal_eval(value.strip())

    def parse_config_file(path: str):
        with open(path) as fh:
            parse_config(fh.read())

    def configurable(cls):
        """Class decorator: constructor keyword arguments default to the parsed bindings."""
        init = cls.__init__

        @functools.wraps(init)
        def bound_init(self, *args, **kwargs):
            merged = dict(_bindings.get(cls.__name__, {}))
            merged.update(kwargs)
            init(self, *args, **merged)

        cls.__init__ = bound_init
        return cls

"""Import stub (TEST INFRASTRUCTURE, oracle/ only): lets the reference's NumPy code path import
without JAX. Every ``isinstance(x, jnp.ndarray)`` in the reference is False with this sentinel,
so the reference takes its NumPy branch. Nothing here computes anything."""
from . import numpy  # noqa: F401


class Array:  # scipy's array-API helper looks up sys.modules['jax'].Array
    pass


class _Random:
    @staticmethod
    def PRNGKey(seed):
        return ("stub-prng-key", int(seed))

    @staticmethod
    def split(key, num=2):
        return tuple((key, i) for i in range(num))


random = _Random()


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn

from . import jax  # noqa

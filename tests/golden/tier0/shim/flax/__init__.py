from . import serialization, linen, metrics  # noqa

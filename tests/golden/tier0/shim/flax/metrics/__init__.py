from . import tensorboard  # noqa

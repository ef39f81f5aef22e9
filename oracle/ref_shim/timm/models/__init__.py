from . import layers, helpers, registry  # noqa: F401

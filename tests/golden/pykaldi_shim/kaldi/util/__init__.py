from . import table  # noqa: F401

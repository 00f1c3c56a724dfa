from . import functions, mel, plp, window  # noqa: F401

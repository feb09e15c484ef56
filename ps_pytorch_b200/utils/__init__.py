from .metrics import accuracy  # noqa: F401

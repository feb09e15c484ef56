from .distributed_backward import backward  # noqa: F401

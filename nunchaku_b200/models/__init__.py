from .linear import SVDQW4A4Linear  # noqa: F401

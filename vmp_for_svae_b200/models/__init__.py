from . import gmm, smm, svae  # noqa: F401

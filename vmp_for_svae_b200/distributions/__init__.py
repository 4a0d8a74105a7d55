from . import dirichlet, gaussian, niw, student_t  # noqa: F401

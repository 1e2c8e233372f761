from jamie_b200.jamie import JAMIE  # noqa: F401

from jamie_b200.model import edModelVar  # noqa: F401

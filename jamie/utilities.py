from jamie_b200.utilities import identity, preclass, time_logger  # noqa: F401

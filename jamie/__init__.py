"""Import-path compatibility package: ``from jamie import JAMIE`` and the module paths that reference checkpoints pickle
by name (``jamie.model.edModelVar``, ``jamie.utilities.preclass``, ``jamie.utilities.identity``) resolve to the
B200-native implementation in ``jamie_b200``."""
from jamie_b200 import __version__  # noqa: F401
from jamie_b200.jamie import JAMIE  # noqa: F401
from . import model, utilities, evaluation  # noqa: F401

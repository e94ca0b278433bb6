"""Device-backed counterparts of qradient.circuit_logic (McClean, Qaoa, MeynardClassifier)."""
from .base import ParametrizedCircuit, progbar_range  # noqa: F401
from .mc_clean import McClean  # noqa: F401
from .qaoa import Qaoa  # noqa: F401
from .meynard_classifier import MeynardClassifier  # noqa: F401

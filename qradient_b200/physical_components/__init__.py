"""Device-backed counterparts of qradient.physical_components (State, Gates, Observable)."""
from .observable import Observable, Projector  # noqa: F401
from .gates import Gates  # noqa: F401
from .state import State  # noqa: F401

"""Host-side mirror of the OpenOpt interface around the de solver (pure Python, no compute)."""
from .log import OpenOptException  # noqa: F401
from .problem import GLP, NLP  # noqa: F401
from .registry import baseSolver, oosolver  # noqa: F401

"""B200-native engine for the CausalImpact fit / posterior-predictive path."""
from ._engine import Engine, EngineError, ProblemSpec  # noqa: F401
from .model import build_problem, initial_theta  # noqa: F401

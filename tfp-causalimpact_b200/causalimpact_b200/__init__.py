"""B200-native engine for the CausalImpact fit / posterior-predictive path.

Drop-in surface (reference causalimpact/__init__.py:29-37):
    import causalimpact_b200 as causalimpact
    ci = causalimpact.fit_causalimpact(data, pre_period, post_period)
"""
__version__ = "0.1.0"

from ._engine import Comm, DeviceArray, Engine, EngineError, ProblemSpec, comm_unique_id  # noqa: F401
from .api import (CausalImpactAnalysis, CausalImpactPosteriorSamples, DataOptions,  # noqa: F401
                  EngineOptions, InferenceOptions, ModelOptions, PanelAnalysis, Seasons,
                  fit_causalimpact, fit_causalimpact_many)
from .frame import CausalImpactData, InputDateType  # noqa: F401
from .model import build_problem, initial_theta  # noqa: F401
from .panel import PanelResult, fit_causalimpact_panel, prepare_panel  # noqa: F401
from .report import plot, summary  # noqa: F401

"""B200-native Sugarscape timestep engine (drop-in for `View.updateGame`).

Only what the hot path needs: csrc/ (CUDA kernels + C-ABI), the ctypes binding
(capi), the settings.json loader (settings) and the host-side mirror of the
reference's step interface (engine, dropin).
"""
from .capi import SsConfig, SsStepStats, STATS_DTYPE, RNG_MT19937, RNG_KEYED, EngineLibraryMissing  # noqa: F401
from .settings import (ConflictingRuleException, MissingRuleException, UnsupportedRuleError,  # noqa: F401
                       config_from_settings, init_params_from_settings, load_settings, rule_check)

__all__ = ["SsConfig", "SsStepStats", "STATS_DTYPE", "RNG_MT19937", "RNG_KEYED", "EngineLibraryMissing",
           "ConflictingRuleException", "MissingRuleException", "UnsupportedRuleError",
           "config_from_settings", "init_params_from_settings", "load_settings", "rule_check"]

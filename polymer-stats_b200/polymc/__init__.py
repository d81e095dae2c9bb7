"""polymc — Python host side of the B200-native fixed-force MCMC hot path of
grasingerm/polymer-stats (mcmc_eap_chain.jl).  All compute goes through libpolymc_b200.so."""
from .lib import (AVG_NAMES, CHAIN_TYPES, CaseTable, ENERGY_TYPES, RESULT_COLS, RESULT_NAMES, Ensemble, MultiEnsemble, PmcCase,
                  PolymcError, build, device_count, fp64_peak_probe, load, make_case, release_cached_memory)

__all__ = ["AVG_NAMES", "CHAIN_TYPES", "CaseTable", "ENERGY_TYPES", "RESULT_COLS", "RESULT_NAMES", "Ensemble", "MultiEnsemble",
           "PmcCase", "PolymcError", "build",
           "device_count", "fp64_peak_probe", "load", "make_case", "release_cached_memory"]

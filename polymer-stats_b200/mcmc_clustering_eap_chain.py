#!/usr/bin/env python
"""Drop-in command line twin of the reference's `julia mcmc_clustering_eap_chain.jl ...` (same options,
same 12-line stdout block, same CSV files), running on the B200 through libpolymc_b200.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from polymc.mcmc_clustering import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())

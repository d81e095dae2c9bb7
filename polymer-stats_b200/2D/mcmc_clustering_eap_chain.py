#!/usr/bin/env python
"""Drop-in command line twin of the reference's `julia 2D/mcmc_clustering_eap_chain.jl ...` (planar chains:
same options, same 10-line stdout block with 2-vectors, same CSV files) on libpolymc_b200.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polymc.mcmc_clustering_2d import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())

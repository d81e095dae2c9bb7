#!/bin/bash
# Wall clock of the drop-in CLI twins used the way the reference's launchers use `julia mcmc_*.jl` (one process per case,
# run/Ising_2025-12-17.jl: 20 cases x 25 runs of n=100, 2.5e6 trials), against one batched study call (run_sweep.py).
cd "$(dirname "$0")/.."
P=polymer-stats_b200
tmp=$(mktemp -d)
t() { python -c "import subprocess,sys,time; s=time.time(); r=subprocess.run(sys.argv[1:],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL); print(f'{time.time()-s:.2f}' + ('' if r.returncode == 0 else f' (rc {r.returncode})'))" "$@"; }
echo "one case per process (the launcher workflow), clustering driver, Ising n=100, bend-mod 0.5, 250000 trials + 5x20000 burn-in:"
echo "  --replicas 1 : $(t python $P/mcmc_clustering_eap_chain.py -n 100 --energy-type Ising --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 --num-steps 250000 --burn-in 20000 -v 0 --prefix $tmp/a --seed 1) s"
echo "  --replicas 25: $(t python $P/mcmc_clustering_eap_chain.py -n 100 --energy-type Ising --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 --num-steps 250000 --burn-in 20000 -v 0 --prefix $tmp/b --seed 1 --replicas 25) s"
echo "same, all-pairs energy:"
echo "  --replicas 1 : $(t python $P/mcmc_clustering_eap_chain.py -n 100 --energy-type interacting --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 --num-steps 250000 --burn-in 20000 -v 0 --prefix $tmp/c --seed 1) s"
echo "  --replicas 25: $(t python $P/mcmc_clustering_eap_chain.py -n 100 --energy-type interacting --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 --num-steps 250000 --burn-in 20000 -v 0 --prefix $tmp/d --seed 1 --replicas 25) s"
echo "python start-up alone (import numpy + the package, no run): $(t python -c 'import sys; sys.path.insert(0, "polymer-stats_b200"); import polymc; polymc.load()') s"
echo "the whole study in one call (20 cases x 25 runs, Ising, same trial counts), run_sweep.py:"
echo "  $(t python $P/run_sweep.py --driver clustering --grid E0=0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8,0.9,1.0,1.1,1.2,1.3,1.4,1.5,1.6,1.7,1.8,1.9,2.0 --runs 25 --out $tmp/study.csv -- -n 100 --energy-type Ising --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 --num-steps 250000 --burn-in 20000) s"
rm -rf $tmp

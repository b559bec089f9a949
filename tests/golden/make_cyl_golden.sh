#!/bin/bash
# Generates tests/golden/cyl/ (run in the build container, where /root/reference exists): the reference's own 4x8
# Heisenberg cylinder input (tests/2Dheisenberg.cpp:292-318, 64 fixture files under tests/2dHeisenberg/) turned into a
# U(1) bMPO by the reference's own helper functions (oracle/ref_cyl.cpp compiles its test translation unit in place),
# coalesced, plus random_bMPS(4, H, Z(0), {}, seed 0); then the reference's dmrg() for 5 sweeps at maximum_bond 48,
# cutoff 1e-12 (convergence criterion 0: every sweep runs). The per-sweep energies go to cyl/reference_sweeps.txt.
set -e
cd "$(dirname "$0")/../.."
make -C oracle ref
mkdir -p tests/golden/cyl
oracle/_ref/ref_cyl tests/golden/cyl 48 1e-12 0.0 5 0 --threads 8 | tee tests/golden/cyl/reference_sweeps.txt

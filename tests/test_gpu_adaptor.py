"""The reference-side C++ binding of INTEGRATION.md, compiled against the real reference headers
(oracle/adaptor_check.cpp -> oracle/_ref/adaptor_check, built by `make -C oracle ref` in the build container; the binary
travels to the GPU box). It converts quantit::btensor <-> the engine's C-ABI handles with the reference's PUBLIC accessors
only, runs the reference entry point and the engine's replacement on the same btensors in one process, and compares:
bit-exact structure, values / singular values / DMRG energies to the north star's tolerance.
Skipped when the compiled reference is absent."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adaptor_check")
LIB = os.path.join(ROOT, "quantit_b200", "libqtb.so")
G = os.path.join(ROOT, "tests", "golden")


def run(*args):
    out = subprocess.run([BIN, "--lib", LIB] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
    print(out.stdout, out.stderr[-1500:])
    assert out.returncode == 0 and "ADAPTOR_OK" in out.stdout, (out.stdout, out.stderr[-1500:])
    return out.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="compiled reference-side adaptor (oracle/_ref) not present")
@pytest.mark.parametrize("case,da,db", [("tdot1", "3", "0"), ("tdot2", "1,3", "2,0"), ("tdot3", "-", "-"), ("tdot4", "2,0,1", "2,0,1")])
def test_adaptor_tensordot(case, da, db):
    """btensor::tensordot through the C++ binding against the reference's own result, same btensor operands"""
    out = run("tdot", os.path.join(G, f"{case}_A.qtbt"), os.path.join(G, f"{case}_B.qtbt"), da, db)
    assert "structure_identical 1" in out


@pytest.mark.skipif(not os.path.exists(BIN), reason="compiled reference-side adaptor (oracle/_ref) not present")
@pytest.mark.parametrize("tol,mn,mx", [(1e-3, 1, 1000), (1e-1, 1, 1000), (0.0, 1, 12)])
def test_adaptor_svd_truncated(tol, mn, mx):
    """svd(A, split, tol, min, max, pow) through the C++ binding: U, d, V structures identical to the reference's,
    singular values and U.d.V^T equal"""
    out = run("svdt", os.path.join(G, "svd_theta.qtbt"), 2, tol, mn, mx, 2)
    assert "structure_identical 1" in out


@pytest.mark.skipif(not os.path.exists(BIN), reason="compiled reference-side adaptor (oracle/_ref) not present")
def test_adaptor_dmrg():
    """quantit::dmrg(bMPO&, bMPS&, options) of the reference and the engine's qtb_dmrg behind the same btensor inputs: the
    converged energies agree to 1e-10 and the engine's final bMPS, read back into the reference, has that energy under the
    REFERENCE's own contract()"""
    run("heis", 10, 40, 1e-10, 1e-9, 20)

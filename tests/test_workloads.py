"""CPU tests of the host-side logic: workload generators and the QTBT exchange format."""
import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl


def test_bond_generator():
    for n, D, s in [(41, 2048, 6.0), (15, 4096, 1.6), (17, 8192, 1.8), (5, 24, 1.0)]:
        sizes, charges = wl.bond(n, D, s, 2)
        assert sum(sizes) == D and len(sizes) == n and min(sizes) >= 1
        assert charges[n // 2] == (0,) and charges[0] == (-2 * (n // 2),)


def test_heisenberg_W_is_the_reference_mpo():
    """dense W equals the reference's Heisenberg_impl tensor (sources/models.cpp:24-57) for J' = 1/4"""
    W = orc.BT(**wl.heisenberg_W())
    dense = W.to_dense()
    sz, sp = np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [0.0, 0.0]])
    assert np.allclose(dense[4, :, 3, :], 0.25 * sz) and np.allclose(dense[3, :, 0, :], sz)
    assert np.allclose(dense[1, :, 0, :], sp) and np.allclose(dense[4, :, 1, :], 0.5 * sp.T)
    # two-site Hamiltonian from the MPO: <s1' s2'| H |s1 s2> = J'(2 S+S- + 2 S-S+ + SzSz)
    H2 = orc.compute_2sites_hamil(W, W).to_dense()[4, :, :, 0, :, :].reshape(4, 4)
    want = 0.25 * (np.kron(sz, sz) + 2 * np.kron(sp, sp.T) + 2 * np.kron(sp.T, sp))
    assert np.allclose(H2, want)


def test_qtbt_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    t = orc.BT(**wl.random_btensor(rng, 3, nc=2))
    p = str(tmp_path / "t.qtbt")
    orc.write_qtbt(t, p)
    r = orc.read_qtbt(p)
    assert orc.same_structure(t, r) and orc.max_rel_err(t, r) == 0.0


def test_chain_generators_feed_a_valid_dmrg():
    """heisenberg_mpo + random_mps (the bench's DMRG inputs, generated without the reference): right-orthonormal MPS of
    the requested charge, and the oracle's two-site DMRG on them finds the exact L=8 open-chain ground energy"""
    import qtb_oracle as orc
    L = 8
    H = [orc.BT(**h) for h in wl.heisenberg_mpo(L)]
    P = [orc.BT(**p) for p in wl.random_mps(L, 4, 0, seed=1)]
    E = orc.dmrg(P, H, 0, 1e-12, 1e-12, 32, 4, 30)
    assert abs(E - (-3.374932598687897)) < 1e-10
    S = wl.random_mps(12, 16, 2, seed=3)
    assert S[-1]["cvals"][2] == [(-2,)] and S[0]["cvals"][0] == [(0,)]
    for k in range(1, 12):
        t = S[k]
        for ql in range(len(t["sec_sizes"][0])):
            keys = sorted(key for key in t["blocks"] if key[0] == ql)
            M = np.concatenate([t["blocks"][key].reshape(t["blocks"][key].shape[0], -1) for key in keys], axis=1)
            assert np.allclose(M @ M.T, np.eye(M.shape[0]), atol=1e-12)


def _dense_from_mpo(sites):
    def dense_site(s):
        D = np.zeros((len(s["sec_sizes"][0]), 2, len(s["sec_sizes"][2]), 2))
        for idx, v in s["blocks"].items():
            D[idx] = v.reshape(())
        return D
    M = dense_site(sites[0])[0].transpose(1, 0, 2)  # [wr, s', s]
    for k in range(1, len(sites)):
        W = dense_site(sites[k])
        M = np.einsum("wab,wcxd->xacbd", M, W).reshape(W.shape[2], M.shape[1] * 2, M.shape[2] * 2)
    return M[0]


def _dense_heisenberg(L, bonds, Jp):
    sz, up, I = np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [0.0, 0.0]]), np.eye(2)
    dn = up.T

    def op(o, i):
        m = np.array([[1.0]])
        for k in range(L):
            m = np.kron(m, o if k == i else I)
        return m
    return sum(Jp * (2 * op(up, i) @ op(dn, j) + 2 * op(dn, i) @ op(up, j) + op(sz, i) @ op(sz, j)) for i, j in bonds)


@pytest.mark.parametrize("Lx,Ly", [(3, 3), (2, 4), (4, 2)])
def test_cylinder_mpo_is_the_heisenberg_hamiltonian(Lx, Ly):
    """the finite-state-machine U(1) MPO of workloads.heisenberg_bonds_mpo (BASELINE.json configs[3] generator) against
    the Hamiltonian built term by term"""
    from quantit_b200 import workloads as wl
    bonds = wl.cylinder_bonds(Lx, Ly)
    assert np.abs(_dense_from_mpo(wl.heisenberg_cylinder_mpo(Lx, Ly)) - _dense_heisenberg(Lx * Ly, bonds, 0.25)).max() == 0.0


def test_bond_list_mpo_reduces_to_the_chain():
    from quantit_b200 import workloads as wl
    chain = wl.heisenberg_bonds_mpo(6, [(i, i + 1) for i in range(5)])
    assert np.abs(_dense_from_mpo(chain) - _dense_from_mpo(wl.heisenberg_mpo(6))).max() == 0.0

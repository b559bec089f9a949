"""numpy model of the batched one-sided block Jacobi (qtb_svd.cu), used to choose the iteration's parameters:
outer sweeps to convergence as a function of block width, inner sweeps per visit, column pre-ordering and
preconditioning.   python profiles/r2/jacobi_model.py [n]"""
import numpy as np, sys, time

def make(m, n, decay, rng):
    u, _ = np.linalg.qr(rng.standard_normal((m, m)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.exp(-decay * np.arange(n))
    return (u[:, :n] * s) @ v.T

def rounds(pe):
    """tournament rounds over pe players"""
    out = []
    for step in range(pe - 1):
        pr = [(pe - 1, step)]
        for k in range(1, pe // 2):
            pr.append(((step + k) % (pe - 1), (step - k) % (pe - 1)))
        out.append(pr)
    return out

def inner_jacobi(G, sweeps, stop_rel=None):
    """round-parallel cyclic two-sided Jacobi on symmetric G; returns J with eigenvalues descending"""
    p = G.shape[0]; pe = p + (p & 1)
    G = G.copy(); J = np.eye(p)
    rr = rounds(pe)
    for sw in range(sweeps):
        gmax = 0.0
        for pr in rr:
            T = np.eye(p)
            for (x, y) in pr:
                if x >= p or y >= p: continue
                gxy = G[x, y]; sc = abs(G[x, x] * G[y, y])
                if gxy == 0 or gxy * gxy <= 1e-34 * sc: continue
                if sc > 0: gmax = max(gmax, gxy * gxy / sc)
                tau = (G[y, y] - G[x, x]) / (2 * gxy)
                t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
                c = 1 / np.sqrt(1 + t * t); s = t * c
                T[x, x] = c; T[y, y] = c; T[x, y] = s; T[y, x] = -s
            G = T.T @ G @ T
            G = (G + G.T) / 2
            J = J @ T
        if gmax < 1e-30: break
        if stop_rel is not None and gmax < 1e-2 and gmax * gmax < stop_rel: break
    o = np.argsort(-np.diag(G), kind="stable")
    return J[:, o]

def block_jacobi(A, b, inner=4, tol=None, maxsweeps=40, adaptive=False):
    m, n = A.shape
    X = A.copy()
    nb = (n + b - 1) // b
    tol = tol or (1e-14 + 4.5e-16 * np.sqrt(m))
    blocks = [np.arange(i * b, min(n, (i + 1) * b)) for i in range(nb)]
    hist = []; visits = []
    pairs = [(i, j) for i in range(nb) for j in range(i + 1, nb)]
    for sweep in range(maxsweeps):
        worst = 0.0; nrot = 0
        for (i, j) in pairs:
            idx = np.concatenate([blocks[i], blocks[j]])
            P = X[:, idx]
            G = P.T @ P
            d = np.diag(G).copy()
            dd = np.outer(d, d)
            S = np.where(dd > 0, np.abs(G) / np.sqrt(np.where(dd > 0, dd, 1)), 0.0); np.fill_diagonal(S, 0)
            g = S.max()
            worst = max(worst, g)
            if g <= tol: continue
            nrot += 1
            J = inner_jacobi(G, inner, stop_rel=(1e-2 * g * g) if adaptive else None)
            X[:, idx] = P @ J
        hist.append(worst); visits.append(nrot)
        if worst < tol: break
    return len(hist), hist, visits, X

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for decay in (0.0, 0.03, 0.15):
        A = make(m, n, decay, rng) if decay > 0 else rng.standard_normal((m, n))
        ref = np.linalg.svd(A, compute_uv=False)
        for b in (16,):
            for pre in ("none", "qrT"):
                for inner in (1, 2, 4, 12):
                    B = A
                    if pre == "sort":
                        B = A[:, np.argsort(-np.linalg.norm(A, axis=0))]
                    elif pre == "qrT":  # Drmac-Veselic: Jacobi on R^T of the column-sorted matrix
                        B = np.linalg.qr(A[:, np.argsort(-np.linalg.norm(A, axis=0))])[1].T.copy()
                    t0 = time.time()
                    ns, hist, visits, X = block_jacobi(B, b, inner)
                    sv = np.sort(np.linalg.norm(X, axis=0))[::-1]
                    print("decay %.2f b %2d pre %-5s inner %2d sweeps %2d visits %5d sv err %.1e (%.0fs) hist %s" % (
                        decay, b, pre, inner, ns, sum(visits), np.abs(sv - ref).max() / ref[0], time.time() - t0,
                        " ".join("%.0e" % h for h in hist)), flush=True)

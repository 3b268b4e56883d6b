"""The solver sweep of the CUDA fast path (csrc/hrl_ant.cuh, `delassus_sweep`) restated in numpy and held to the
velocity-space sweep it replaces (the same projected Gauss-Seidel the oracle runs, oracle/hrl_oracle.c `solve_rows`;
Bullet's btMultiBodyConstraintSolver row order [3P-MEM]).

Velocity space: every visit takes a 14-term dot of the whitened row with the accumulated velocity change, clamps the
impulse and applies `a * dl` to the velocity - a dependent chain dot -> clamp -> axpy per visit.
Delassus space: with What[i][j] = (a_i . a_j) / (a_j . a_j) (zero diagonal) the sweep tracks, per row, the UNCLAMPED
target t_j = lambda_j + rhs'_j - (a_j . dv) / (a_j . a_j); a visit is clamp(t_i) -> dl -> t_j -= What[i][j] dl for all j:
independent FMAs, no dot on the chain.  Same iterates in exact arithmetic; the final velocity change is sum_j lambda_j a_j."""
import numpy as np


def sweep_velocity(A, rhs, kind, mu, iters, max_imp):
    """A: [R,14] whitened rows; rhs: [R] (already divided by a.a); kind: 0 limit, 1 normal, 2/3 friction t1/t2 of the
    contact whose normal row index is in `mu[r][1]`.  Returns (dv, lam)."""
    R = len(A)
    dinv = np.array([1.0 / (a @ a) if a @ a > 0 else 0.0 for a in A])
    lam = np.zeros(R)
    dv = np.zeros(14)
    lim = [r for r in range(R) if kind[r] == 0]
    nrm = [r for r in range(R) if kind[r] == 1]
    fri = [r for r in range(R) if kind[r] == 2]
    for it in range(iters):
        order = lim if (it & 1) else lim[::-1]
        for r in order:
            nl = min(max(lam[r] + rhs[r] - (A[r] @ dv) * dinv[r], 0.0), max_imp)
            dv += A[r] * (nl - lam[r]); lam[r] = nl
        for r in nrm:
            nl = max(lam[r] + rhs[r] - (A[r] @ dv) * dinv[r], 0.0)
            dv += A[r] * (nl - lam[r]); lam[r] = nl
        for r in fri:
            a, b = r, r + 1
            n = int(mu[r][1])
            na = lam[a] + rhs[a] - (A[a] @ dv) * dinv[a]
            nb = lam[b] + rhs[b] - (A[b] @ dv) * dinv[b]
            limv = mu[r][0] * lam[n]
            l2 = na * na + nb * nb
            sc = limv / np.sqrt(max(l2, 1e-30)) if l2 > limv * limv else 1.0
            if lam[n] > 0:
                na, nb = na * sc, nb * sc
            else:
                na, nb = lam[a], lam[b]
            dv += A[a] * (na - lam[a]) + A[b] * (nb - lam[b])
            lam[a], lam[b] = na, nb
    return dv, lam


def sweep_delassus(A, rhs, kind, mu, iters, max_imp):
    R = len(A)
    dinv = np.array([1.0 / (a @ a) if a @ a > 0 else 0.0 for a in A])
    What = (A @ A.T) * dinv[None, :]
    np.fill_diagonal(What, 0.0)
    lam = np.zeros(R)
    t = rhs.copy()
    lim = [r for r in range(R) if kind[r] == 0]
    nrm = [r for r in range(R) if kind[r] == 1]
    fri = [r for r in range(R) if kind[r] == 2]
    for it in range(iters):
        order = lim if (it & 1) else lim[::-1]
        for r in order:
            nl = min(max(t[r], 0.0), max_imp)
            dl = nl - lam[r]; lam[r] = nl
            t -= What[r] * dl
        for r in nrm:
            nl = max(t[r], 0.0)
            dl = nl - lam[r]; lam[r] = nl
            t -= What[r] * dl
        for r in fri:
            a, b = r, r + 1
            n = int(mu[r][1])
            na, nb = t[a], t[b]
            limv = mu[r][0] * lam[n]
            l2 = na * na + nb * nb
            sc = limv / np.sqrt(max(l2, 1e-30)) if l2 > limv * limv else 1.0
            if lam[n] > 0:
                na, nb = na * sc, nb * sc
            else:
                na, nb = lam[a], lam[b]
            dla, dlb = na - lam[a], nb - lam[b]
            lam[a], lam[b] = na, nb
            t -= What[a] * dla + What[b] * dlb
    return A.T @ lam, lam


def random_problem(rng, n_lim, n_con):
    rows, rhs, kind, mu = [], [], [], []
    def row(leg):
        a = np.zeros(14)
        a[:6] = rng.normal(size=6) * 0.3
        a[6 + 2 * leg: 8 + 2 * leg] = rng.normal(size=2)
        return a
    for _ in range(n_lim):
        rows.append(row(rng.integers(4))); rhs.append(rng.normal() * 2); kind.append(0); mu.append((0.0, -1))
    nidx = []
    for _ in range(n_con):
        leg = rng.integers(4)
        nidx.append((len(rows), leg))
        rows.append(row(leg)); rhs.append(rng.normal() * 2 + 1); kind.append(1); mu.append((0.0, -1))
    for n, leg in nidx:
        rows.append(row(leg)); rhs.append(rng.normal()); kind.append(2); mu.append((0.8, n))
        rows.append(row(leg)); rhs.append(rng.normal()); kind.append(3); mu.append((0.8, n))
    return np.array(rows), np.array(rhs), kind, mu


def test_delassus_sweep_equals_velocity_sweep():
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(200):
        A, rhs, kind, mu = random_problem(rng, int(rng.integers(0, 9)), int(rng.integers(0, 5)))
        if len(A) == 0:
            continue
        dv0, l0 = sweep_velocity(A, rhs, kind, mu, 5, 50.0)
        dv1, l1 = sweep_delassus(A, rhs, kind, mu, 5, 50.0)
        worst = max(worst, np.abs(dv0 - dv1).max(), np.abs(l0 - l1).max())
    assert worst < 1e-9, worst

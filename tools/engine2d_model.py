"""numpy model of the 2-D cyclic group engine (csrc/local_step_2d.cuh): 16 lanes (pr, pc) of a 4x4 grid own
A(i, c) with i = 4a+pr, c = 4b+pc (register Areg[a][b], b <= a).  Right-looking Cholesky with deferred scaling:
step j publishes the UNSCALED pivot column (zeros for finished rows), every lane reads the entries of its rows
(multipliers m = s * inv^2) and of its columns, and updates Areg[a][b] -= m[a] * cv[b]; the zeros keep finished
rows / columns untouched, so all loop bounds are lane-independent.  Forward substitution rides along on a RHS
distributed as g[t] <-> row 4(4t+pc)+pr; back substitution goes by 4-column blocks with partial sums per lane.
Validates the index logic on the CPU before it is written in CUDA."""
import numpy as np


def run(D, seed=0):
    NB = D // 4
    rs = np.random.RandomState(seed)
    M = rs.randn(D, D); P2 = M @ M.T + D * np.eye(D)
    p1 = np.logaddexp(0, rs.randn(D)); g = rs.randn(D); g1 = rs.randn(D); eps = rs.randn(D)
    Pt = P2 + np.diag(p1)
    lanes = [(pr, pc) for pr in range(4) for pc in range(4)]
    A = {ln: np.zeros((NB, NB)) for ln in lanes}
    for (pr, pc) in lanes:
        for a in range(NB):
            for b in range(a + 1):
                A[(pr, pc)][a, b] = Pt[4 * a + pr, 4 * b + pc]         # includes garbage where 4a+pr < 4b+pc (a == b)
    NT = NB // 4 if NB >= 4 else 1
    # RHS layout: lane (pr,pc) holds rows i = 4a+pr for a = 4t+pc, t < NB/4   (NB multiple of 4)
    assert NB % 4 == 0
    G = {ln: np.array([[g[4 * (4 * t + ln[1]) + ln[0]], g1[4 * (4 * t + ln[1]) + ln[0]]] for t in range(NB // 4)]) for ln in lanes}
    avec = np.zeros(D); idiag = np.zeros(D)
    myinv = {ln: 0.0 for ln in lanes}
    q = 0.0; hl = 0.0
    for j in range(D):
        ja, jr = divmod(j, 4)
        colbuf = np.zeros((4, NB))
        # publish: lanes with pc == jr
        for (pr, pc) in lanes:
            if pc == jr:
                for a in range(ja, NB):
                    colbuf[pr, a] = A[(pr, pc)][a, ja] if 4 * a + pr > j else 0.0
        piv = A[(jr, jr)][ja, ja]
        own_g = (jr, ja % 4)
        gj, g1j = G[own_g][ja // 4]
        inv = 1.0 / np.sqrt(piv); inv2 = inv * inv
        hl += np.log(piv); q += gj * g1j * inv2
        avec[j] = gj * inv; idiag[j] = inv
        for ln in lanes:
            pr, pc = ln
            if pc == jr:
                myinv[ln] = inv
            m = colbuf[pr] * inv2                       # multipliers for own rows (index a)
            cv = colbuf[pc]                             # column values for own columns (index b)
            for a in range(ja, NB):
                for b in range(ja, a + 1):
                    A[ln][a, b] -= m[a] * cv[b]
            for t in range(NB // 4):
                a = 4 * t + pc
                if a >= ja:
                    G[ln][t, 0] -= m[a] * gj
                    G[ln][t, 1] -= m[a] * g1j
        if jr == 3:                                     # block of 4 columns done: every lane scales its column
            for ln in lanes:
                for a in range(ja, NB):
                    A[ln][a, ja] *= myinv[ln]
    L = np.linalg.cholesky(Pt)
    Lm = np.zeros((D, D))
    for (pr, pc) in lanes:
        for a in range(NB):
            for b in range(a + 1):
                i, c = 4 * a + pr, 4 * b + pc
                if i >= c:
                    Lm[i, c] = A[(pr, pc)][a, b]
    assert np.allclose(Lm, L, rtol=1e-9, atol=1e-9), np.abs(Lm - L).max()
    a_ref = np.linalg.solve(L, g); a1_ref = np.linalg.solve(L, g1)
    assert np.allclose(avec, a_ref) and np.isclose(q, a_ref @ a1_ref) and np.isclose(0.5 * hl, np.log(np.diag(L)).sum())
    # ---- back substitution y = L^-T (eps - a): partial sums wpart[b] for c = 4b+pc on every lane
    w = eps - avec
    wpart = {ln: np.array([w[4 * b + ln[1]] if ln[0] == 0 else 0.0 for b in range(NB)]) for ln in lanes}
    y = np.zeros(D)
    for bb in range(NB - 1, -1, -1):
        T = np.array([sum(wpart[(pr, pc)][bb] for pr in range(4)) for pc in range(4)])      # reduce over pr
        yb = np.zeros(4)
        for cc in range(3, -1, -1):                      # 4x4 diagonal block, transposed solve
            yb[cc] = T[cc] * idiag[4 * bb + cc]
            for c2 in range(cc):
                T[c2] -= A[(cc, c2)][bb, bb] * yb[cc]    # L[4bb+cc][4bb+c2] lives on lane (cc, c2)
        y[4 * bb:4 * bb + 4] = yb
        for ln in lanes:
            pr, pc = ln
            for b2 in range(bb):
                wpart[ln][b2] -= A[ln][bb, b2] * yb[pr]  # L[4bb+pr][4b2+pc] * y[4bb+pr]
    y_ref = np.linalg.solve(L.T, eps - a_ref)
    assert np.allclose(y, y_ref, rtol=1e-9, atol=1e-9), np.abs(y - y_ref).max()
    return True


if __name__ == '__main__':
    for D in (16, 32, 48, 64):
        assert run(D, seed=D)
        print('ok', D)

"""numpy model of the 2-D cyclic group engine (csrc/local_step_fast2d.cuh), lane by lane, to validate the index
logic on the CPU before it is written in CUDA.

16 lanes (pr, pc) of a 4x4 grid own A(i, c) with i = 4a+pr, c = 4b+pc as register Areg[a][b], b <= a < NB, plus an
augmented row a = NB that carries the two right-hand sides (lanes pr = 0: g = P2 d, pr = 1: g1 = P1 d, pr >= 2: 0), so
the forward substitutions are part of the factorisation.  Right-looking Cholesky with DEFERRED scaling: step j
publishes the UNSCALED pivot column (zero for finished rows), every lane reads the entries of its row class
(multipliers m = s * inv^2) and of its column class (cv = s) and updates Areg[a][b] -= m[a] * cv[b]; the zeros keep
finished rows untouched, so all loop bounds are lane-independent.  Columns are scaled by their 1/L_jj when their block
of four is done.  g = P2 d is a symmetric mat-vec on the lower-triangular registers (column sums reduced over pr, row
sums over pc); back substitution goes by 4-column blocks with per-lane partial sums and shuffles only; the quadratic
form |W (x - m)|^2 is a row-class partial mat-vec reduced over pc."""
import numpy as np


def run(D, seed=0):
    NB = D // 4
    rs = np.random.RandomState(seed)
    M = rs.randn(D, D); P2 = M @ M.T + D * np.eye(D)
    p1 = np.logaddexp(0, rs.randn(D)); d = rs.randn(D); eps = rs.randn(D)
    Wt = np.tril(rs.randn(D, D)); mth = rs.randn(D); mu1 = rs.randn(D)
    Pt = P2 + np.diag(p1)
    lanes = [(pr, pc) for pr in range(4) for pc in range(4)]
    # staged lane-major records: lower triangle, zeros in the upper part of the diagonal blocks
    E = {ln: np.zeros((NB, NB)) for ln in lanes}
    Wl = {ln: np.zeros((NB, NB)) for ln in lanes}
    for (pr, pc) in lanes:
        for a in range(NB):
            for b in range(a + 1):
                i, c = 4 * a + pr, 4 * b + pc
                if i >= c:
                    E[(pr, pc)][a, b] = P2[i, c]
                    Wl[(pr, pc)][a, b] = Wt[i, c]
    # ---- phase 1: symmetric mat-vec g = P2 d on the lower-triangular registers
    colsum = {ln: np.zeros(NB) for ln in lanes}
    rowsum = {ln: np.zeros(NB) for ln in lanes}
    for ln in lanes:
        pr, pc = ln
        for a in range(NB):
            for b in range(a + 1):
                e = E[ln][a, b]
                colsum[ln][b] += e * d[4 * a + pr]                              # P2[i][c] d[i]  -> g[c]
                if not (a == b and pr == pc):
                    rowsum[ln][a] += e * d[4 * b + pc]                          # P2[i][c] d[c]  -> g[i]   (i > c)
    colred = {ln: sum(colsum[(q, ln[1])] for q in range(4)) for ln in lanes}   # reduce over pr (xor 4, 8)
    rowred = {ln: sum(rowsum[(ln[0], q)] for q in range(4)) for ln in lanes}   # reduce over pc (xor 1, 2)
    A = {ln: np.zeros((NB + 1, NB)) for ln in lanes}
    for ln in lanes:
        pr, pc = ln
        A[ln][:NB] = E[ln]
        if pr == pc:
            for a in range(NB):
                A[ln][a, a] += p1[4 * a + pr]
        for b in range(NB):
            gfull = colred[ln][b] + rowred[(pc, 0)][b]      # row index 4b+pc lives on lanes with pr' = pc (shfl from lane (pc, 0))
            A[ln][NB, b] = gfull if pr == 0 else (p1[4 * b + pc] * d[4 * b + pc] if pr == 1 else 0.0)
    assert np.allclose([A[(0, c % 4)][NB, c // 4] for c in range(D)], P2 @ d)
    # ---- phase 2: factorisation with the right-hand sides riding along as row NB
    avec = np.zeros(D); idiag = np.zeros(D)
    myinv = {ln: 0.0 for ln in lanes}
    q = 0.0; hl = 0.0
    for j in range(D):
        ja, jr = divmod(j, 4)
        colbuf = np.full((4, NB + 1), np.nan)
        for (pr, pc) in lanes:
            if pc == jr:
                colbuf[pr, ja] = A[(pr, pc)][ja, ja] if pr > jr else 0.0
                for a in range(ja + 1, NB + 1):
                    colbuf[pr, a] = A[(pr, pc)][a, ja]
        piv = A[(jr, jr)][ja, ja]
        gj, g1j = A[(0, jr)][NB, ja], A[(1, jr)][NB, ja]
        inv = 1.0 / np.sqrt(piv); inv2 = inv * inv
        hl += np.log(piv); q += gj * g1j * inv2
        avec[j] = gj * inv; idiag[j] = inv
        for ln in lanes:
            pr, pc = ln
            if pc == jr:
                myinv[ln] = inv
            m = colbuf[pr] * inv2
            cv = colbuf[pc]
            for a in range(ja, NB + 1):
                for b in range(ja, min(a, NB - 1) + 1):
                    A[ln][a, b] -= m[a] * cv[b]
        if jr == 3:
            for ln in lanes:
                for a in range(ja, NB + 1):
                    A[ln][a, ja] *= myinv[ln]
    L = np.linalg.cholesky(Pt)
    for (pr, pc) in lanes:
        for a in range(NB):
            for b in range(a + 1):
                i, c = 4 * a + pr, 4 * b + pc
                if i >= c:
                    assert abs(A[(pr, pc)][a, b] - L[i, c]) < 1e-9
    a_ref = np.linalg.solve(L, P2 @ d); a1_ref = np.linalg.solve(L, p1 * d)
    assert np.allclose(avec, a_ref) and np.isclose(q, a_ref @ a1_ref) and np.isclose(0.5 * hl, np.log(np.diag(L)).sum())
    assert np.allclose([A[(0, c % 4)][NB, c // 4] for c in range(D)], a_ref)       # scaled row NB = a on lanes pr == 0
    # ---- phase 3: back substitution y = L^-T (eps - a) with per-lane partial sums
    wpart = {ln: np.array([eps[4 * b + ln[1]] - A[ln][NB, b] if ln[0] == 0 else 0.0 for b in range(NB)]) for ln in lanes}
    y = np.zeros(D)
    for bb in range(NB - 1, -1, -1):
        T = np.array([sum(wpart[(pr, pc)][bb] for pr in range(4)) for pc in range(4)])      # reduce over pr
        yb = np.zeros(4)
        for cc in range(3, -1, -1):
            yb[cc] = T[cc] * idiag[4 * bb + cc]
            for c2 in range(cc):
                T[c2] -= A[(cc, c2)][bb, bb] * yb[cc]    # L[4bb+cc][4bb+c2] lives on lane (cc, c2)
        y[4 * bb:4 * bb + 4] = yb
        for ln in lanes:
            pr, pc = ln
            for b2 in range(bb):
                wpart[ln][b2] -= A[ln][bb, b2] * yb[pr]
    assert np.allclose(y, np.linalg.solve(L.T, eps - a_ref), rtol=1e-9, atol=1e-9)
    # ---- quadratic form |W (x - m)|^2
    xm = mu1 + y - mth
    tpart = {ln: np.array([sum(Wl[ln][a, b] * xm[4 * b + ln[1]] for b in range(a + 1)) for a in range(NB)]) for ln in lanes}
    tred = {ln: sum(tpart[(ln[0], qq)] for qq in range(4)) for ln in lanes}         # reduce over pc
    q2 = sum((tred[(pr, 0)] ** 2).sum() for pr in range(4))
    assert np.isclose(q2, ((Wt @ xm) ** 2).sum())
    return True


if __name__ == '__main__':
    for D in (16, 32, 48, 64):
        assert run(D, seed=D)
        print('ok', D)

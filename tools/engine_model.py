"""numpy model of the lane-level algorithm of the group engine (csrc/group_engine.cuh): BS lanes own rows
row(r, lane) = r*BS + lane; A[r][c] slots with garbage above the diagonal; right-looking Cholesky with the column
broadcast, fused 2-RHS forward solve, blocked-transpose back substitution.  Validates index logic on the CPU."""
import numpy as np


def run(D, BS, seed=0):
    ROWS = D // BS
    rs = np.random.RandomState(seed)
    M = rs.randn(D, D); P2 = M @ M.T + D * np.eye(D)
    p1 = np.logaddexp(0, rs.randn(D)); g = rs.randn(D); g1 = rs.randn(D); eps = rs.randn(D)
    lanes = np.arange(BS)
    row = lambda r: r * BS + lanes
    # registers: A[r][c][lane]
    A = np.full((ROWS, D, BS), np.nan)
    for r in range(ROWS):
        for c in range((r + 1) * BS):
            A[r, c] = P2[row(r), c]            # includes garbage region c > row (finite values)
    G = np.stack([g[row(r)] for r in range(ROWS)]); G1 = np.stack([g1[row(r)] for r in range(ROWS)])
    P1 = np.stack([p1[row(r)] for r in range(ROWS)])
    a = np.zeros((ROWS, BS)); idiag = np.zeros((ROWS, BS)); hld = 0.0; q = 0.0
    col = np.zeros((2, D))
    for j in range(D):
        rj, lj = divmod(j, BS)
        dj = A[rj, j] + P1[rj]
        piv = dj[lj]
        inv = 1.0 / np.sqrt(piv); hld += np.log(piv)
        yj = (G[rj] * inv)[lj]; y1j = (G1[rj] * inv)[lj]
        q += yj * y1j
        a[rj] = np.where(lanes == lj, yj, a[rj]); idiag[rj] = np.where(lanes == lj, inv, idiag[rj])
        for r in range(rj, ROWS):
            A[r, j] = np.where((r == rj) & (lanes == lj), piv * inv, A[r, j] * inv)
            col[j & 1, row(r)] = A[r, j]
            G[r] = G[r] - A[r, j] * yj; G1[r] = G1[r] - A[r, j] * y1j
        for r in range(rj, ROWS):
            for c in range(j + 1, (r + 1) * BS):
                A[r, c] = A[r, c] - A[r, j] * col[j & 1, c]
    L = np.linalg.cholesky(P2 + np.diag(p1))
    Lm = np.zeros((D, D))
    for r in range(ROWS):
        for c in range((r + 1) * BS):
            for l in range(BS):
                if c <= r * BS + l:
                    Lm[r * BS + l, c] = A[r, c, l]
    assert np.allclose(Lm, L, rtol=1e-10, atol=1e-10), np.abs(Lm - L).max()
    a_ref = np.linalg.solve(L, g); a1_ref = np.linalg.solve(L, g1)
    assert np.allclose(np.concatenate(list(a)), a_ref) and np.isclose(q, a_ref @ a1_ref)
    assert np.isclose(0.5 * hld, np.log(np.diag(L)).sum())
    # back substitution  y = L^-T (eps - a), blocked transposes
    w = np.stack([eps[row(r)] for r in range(ROWS)]) - a
    y = np.zeros((ROWS, BS))
    for rb in range(ROWS - 1, -1, -1):
        # transposed diagonal block: T[i'][lane] = L[rb*BS+i'][rb*BS+lane]
        tbuf = np.stack([A[rb, rb * BS + cc] for cc in range(BS)], axis=1)    # tbuf[lane_row][cc]
        T = np.stack([tbuf[i, lanes] for i in range(BS)])                      # T[i'][lane]
        for i in range(BS - 1, -1, -1):
            yi = (w[rb] * idiag[rb])[i]
            y[rb] = np.where(lanes == i, yi, y[rb])
            w[rb] = w[rb] - T[i] * yi
        ybuf = y[rb].copy()                                                    # broadcast of y_b
        for rb2 in range(rb):
            tbuf = np.stack([A[rb, rb2 * BS + cc] for cc in range(BS)], axis=1)
            T = np.stack([tbuf[i, lanes] for i in range(BS)])
            for i in range(BS):
                w[rb2] = w[rb2] - T[i] * ybuf[i]
    y_ref = np.linalg.solve(L.T, eps - a_ref)
    assert np.allclose(np.concatenate(list(y)), y_ref, rtol=1e-9, atol=1e-9), np.abs(np.concatenate(list(y)) - y_ref).max()
    return True


if __name__ == '__main__':
    for D, BS in ((64, 16), (64, 32), (32, 8), (32, 16), (16, 4), (48, 16), (16, 16), (8, 8), (24, 8)):
        assert run(D, BS, seed=D + BS)
        print('ok', D, BS)

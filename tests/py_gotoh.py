"""Second, independently written Gotoh implementation (pure Python, three full matrices).

Used only to cross-check the C oracle's scores and end cells on small inputs: it is written
from the textbook recurrence with explicit H/E/F tables and picks the end cell by sorting
candidates, so it shares no control flow with oracle/gotoh_oracle.c.
"""
NEG = -(10 ** 9)


def gotoh(q, r, table, mapper, mode="nw", open=0, gap=0, s1_beg=True, s1_end=True, s2_beg=True, s2_end=True):
    n, m = len(q), len(r)
    if mode != "sg":
        s1_beg = s1_end = s2_beg = s2_end = False
    H = [[0] * (m + 1) for _ in range(n + 1)]
    E = [[NEG] * (m + 1) for _ in range(n + 1)]
    F = [[NEG] * (m + 1) for _ in range(n + 1)]
    for j in range(1, m + 1):
        H[0][j] = 0 if (mode == "sw" or s1_beg) else -open - (j - 1) * gap
    for i in range(1, n + 1):
        H[i][0] = 0 if (mode == "sw" or s2_beg) else -open - (i - 1) * gap
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            E[i][j] = max(E[i][j - 1] - gap, H[i][j - 1] - open)
            F[i][j] = max(F[i - 1][j] - gap, H[i - 1][j] - open)
            s = int(table[mapper[q[i - 1]]][mapper[r[j - 1]]])
            h = max(H[i - 1][j - 1] + s, E[i][j], F[i][j])
            if mode == "sw":
                h = max(h, 0)
            H[i][j] = h
    if mode == "nw" or (mode == "sg" and not s1_end and not s2_end):
        return H[n][m], n - 1, m - 1
    if mode == "sw":
        cands = [(-H[i][j], j - 1, i - 1) for i in range(1, n + 1) for j in range(1, m + 1)]
        best = min(cands)
        if best[0] == 0:
            return 0, 0, 0
        return -best[0], best[2], best[1]
    row = [(-H[n][j], j - 1) for j in range(1, m + 1)] if s1_end else []
    col = [(-H[i][m], i - 1) for i in range(1, n + 1)] if s2_end else []
    if row and col:
        rb, cb = min(row), min(col)
        if -cb[0] > -rb[0]:
            return -cb[0], cb[1], m - 1
        return -rb[0], n - 1, rb[1]
    if row:
        rb = min(row)
        return -rb[0], n - 1, rb[1]
    cb = min(col)
    return -cb[0], cb[1], m - 1

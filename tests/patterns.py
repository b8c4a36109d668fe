"""Small pure-numpy sparsity-pattern builders shared by the tests (no product code)."""
import numpy as np


def grid_pattern(nx, ny, nz, nnc_planes=0):
    """7-point pattern in natural ordering cell = i + nx*(j + ny*k), columns ascending.

    nnc_planes > 0 adds fault-like non-neighbour connections: across the plane i0 = nx//2 cell
    (i0-1, j, k) is also connected to (i0, j, k+s) for s = 1..nnc_planes (structurally symmetric).
    """
    def cid(i, j, k):
        return i + nx * (j + ny * k)
    nbrs = [set() for _ in range(nx * ny * nz)]
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = cid(i, j, k)
                nbrs[c].add(c)
                for (di, dj, dk) in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                    ii, jj, kk = i + di, j + dj, k + dk
                    if ii < nx and jj < ny and kk < nz:
                        d = cid(ii, jj, kk)
                        nbrs[c].add(d)
                        nbrs[d].add(c)
    if nnc_planes > 0 and nx >= 2:
        i0 = nx // 2
        for s in range(1, nnc_planes + 1):
            for k in range(nz - s):
                for j in range(ny):
                    a, b = cid(i0 - 1, j, k), cid(i0, j, k + s)
                    nbrs[a].add(b)
                    nbrs[b].add(a)
    rows = [0]
    cols = []
    for s in nbrs:
        cols.extend(sorted(s))
        rows.append(len(cols))
    return np.array(rows, np.int32), np.array(cols, np.int32)

"""Shared test helpers (dense restatements with numpy for tiny cases)."""
import numpy as np


def oracle_wells(w):
    from oracle import oracle
    if w is None:
        return None
    return oracle.Wells(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv)


def bridge_wells(w, mode="b200"):
    from opm_autodiff_b200 import bridge
    if w is None:
        return bridge.WellContributions(mode, False)
    return bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv, mode)


def oracle_mswells(ms):
    from oracle import oracle
    if not ms:
        return None
    return oracle.MSWells([oracle.MultisegmentWell(m.Mb, m.Bvalues, m.BcolIndices, m.BrowPointers, m.DnumBlocks, m.Dvalues,
                                                   m.DcolPointers, m.DrowIndices, m.Cvalues) for m in ms])


def add_bridge_mswells(wc, ms):
    """Append multisegment wells to a bridge.WellContributions (addMultisegmentWellContribution per well)."""
    for m in ms or ():
        wc.addMultisegmentWellContribution(3, 4, m.Mb, m.Bvalues, m.BcolIndices, m.BrowPointers, m.DnumBlocks, m.Dvalues,
                                           m.DcolPointers, m.DrowIndices, m.Cvalues)
    return wc


def dense_mswell_operator(ms, Nb):
    """sum over multisegment wells of C^T D^-1 B as a dense (3Nb x 3Nb) matrix."""
    Mop = np.zeros((3 * Nb, 3 * Nb))
    for m in ms or ():
        M = 4 * m.Mb
        Bd = np.zeros((M, 3 * Nb))
        Cd = np.zeros((M, 3 * Nb))
        for r in range(m.Mb):
            for blk in range(int(m.BrowPointers[r]), int(m.BrowPointers[r + 1])):
                c = int(m.BcolIndices[blk])
                Bd[4 * r:4 * r + 4, 3 * c:3 * c + 3] += np.asarray(m.Bvalues[blk]).reshape(4, 3)
                Cd[4 * r:4 * r + 4, 3 * c:3 * c + 3] += np.asarray(m.Cvalues[blk]).reshape(4, 3)
        Mop += Cd.T @ np.linalg.solve(m.dense_D(), Bd)
    return Mop


def dense_from_bsr(rows, cols, vals):
    Nb = len(rows) - 1
    A = np.zeros((3 * Nb, 3 * Nb))
    vals = np.asarray(vals).reshape(-1, 3, 3)
    for i in range(Nb):
        for k in range(rows[i], rows[i + 1]):
            A[3 * i:3 * i + 3, 3 * cols[k]:3 * cols[k] + 3] = vals[k]
    return A


def dense_well_operator(w, Nb):
    """C^T D^-1 B as a dense (3Nb x 3Nb) matrix."""
    M = np.zeros((3 * Nb, 3 * Nb))
    if w is None:
        return M
    for i in range(len(w.val_pointers) - 1):
        s, e = int(w.val_pointers[i]), int(w.val_pointers[i + 1])
        Bd = np.zeros((4, 3 * Nb))
        Cd = np.zeros((4, 3 * Nb))
        for p in range(s, e):
            Bd[:, 3 * w.Bcols[p]:3 * w.Bcols[p] + 3] += np.asarray(w.B[p]).reshape(4, 3)
            Cd[:, 3 * w.Ccols[p]:3 * w.Ccols[p] + 3] += np.asarray(w.C[p]).reshape(4, 3)
        M += Cd.T @ np.asarray(w.Dinv[i]).reshape(4, 4) @ Bd
    return M


def relerr(a, b):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def permute_system(rows, cols, vals, b, to, fr, wells=None):
    """P A P^T, P b (and the wells' columns) for an ordering to[natural row] = position, fr[position] = natural row, columns of
    every row sorted ascending -- what reorderBlockedMatrixByPattern / reorderBlockedVectorByPattern do (Reorder.cpp:179-233)."""
    import copy
    rows, cols = np.asarray(rows), np.asarray(cols)
    vals = np.asarray(vals).reshape(-1, 3, 3)
    Nb = len(rows) - 1
    prow = np.zeros(Nb + 1, np.int32)
    pcol = np.zeros(len(cols), np.int32)
    pval = np.zeros_like(vals)
    out = 0
    for p in range(Nb):
        r = fr[p]
        k = np.arange(rows[r], rows[r + 1])
        c = np.asarray(to)[cols[k]]
        o = np.argsort(c, kind="stable")
        n = len(k)
        pcol[out:out + n] = c[o]
        pval[out:out + n] = vals[k[o]]
        out += n
        prow[p + 1] = out
    pb = np.asarray(b).reshape(-1, 3)[np.asarray(fr)].reshape(-1)
    pw = None
    if wells is not None:
        pw = copy.copy(wells)
        pw.Bcols = np.asarray(to)[np.asarray(wells.Bcols)].astype(np.int32)
        pw.Ccols = np.asarray(to)[np.asarray(wells.Ccols)].astype(np.int32)
    return prow, pcol, pval.reshape(-1), pb, pw

"""TEST INFRASTRUCTURE -- numpy restatement of the deterministic background-mode partition.

The reference partitions a column with ``MiniBatchKMeans(n_clusters=k)`` on the projections of the
zero-mean pixels onto ``eig(cov)[1][:, :pcadim]`` (cmf/robust_mf.py:308-313).  That k-means is unseeded, so
the reference's own partition is not reproducible ("parity unpinned" for this step): what can
be pinned is everything downstream of the labels (tests/golden/modes_*.npz use the labels of a seeded
reference run) and the rule the CUDA path uses instead, restated here from
``srcfinder_b200/csrc/k_cluster.cu``.  Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np

INIT_BINS = 4096


def pca_projections(x_nd, pcadim):
    """Projections of the mean-removed rows on the ``pcadim`` leading eigenvectors of ``cov(x)`` (descending
    eigenvalue), each eigenvector signed so that v . mean >= 0 (cmf/robust_mf.py:309-311)."""
    mu = x_nd.mean(axis=0)
    xc = x_nd - mu
    evals, evecs = np.linalg.eigh(np.cov(xc.T, ddof=1))
    order = np.argsort(-evals, kind="stable")[:pcadim]
    v = evecs[:, order]
    v = v * np.where(v.T.dot(mu) < 0.0, -1.0, 1.0)[None, :]
    return xc.dot(v), v


def kmeans_labels(y_np, k, max_iter=100):
    """Lloyd iterations on the quantised projections until no pixel changes cluster, at most ``n / 256`` pixels
    change (that reassignment is kept) or ``max_iter`` passes; returns (labels (n,), reassignment passes)."""
    y = np.asarray(y_np, dtype=np.float64)
    n, pd = y.shape
    mx = float(np.max(np.abs(y))) if n else 0.0
    e = int(np.frexp(mx)[1]) if mx > 0.0 else 0
    scale = float(np.ldexp(1.0, 24 - e))
    q = np.rint(y * scale).astype(np.int64)
    # initial partition: approximately equal-count slices of component 1, from a 4096-bin histogram of it
    lo, hi = int(q[:, 0].min()), int(q[:, 0].max())
    h = ((q[:, 0] - lo) * INIT_BINS) // (hi - lo + 1)
    below = np.concatenate([[0], np.cumsum(np.bincount(h, minlength=INIT_BINS))[:-1]])     # pixels in lower bins
    lab = np.minimum(k - 1, (below[h] * k) // n)
    cen = np.zeros((k, pd))
    qd = q.astype(np.float64)
    it = 0
    while True:
        for c in range(k):
            m = lab == c
            cnt = int(m.sum())
            if cnt > 0:
                cen[c] = q[m].sum(axis=0).astype(np.float64) / float(cnt)
        if it >= max_iter:
            break
        dist = np.zeros((n, k))
        for c in range(k):
            acc = np.zeros(n)
            for p in range(pd):
                df = qd[:, p] - cen[c, p]
                acc = acc + df * df
            dist[:, c] = acc
        new = np.argmin(dist, axis=1)
        nchg = int((new != lab).sum())
        if nchg == 0:
            break
        lab = new
        it += 1
        if nchg * 256 <= n:           # converged to within 2^-8 of the column: keep the reassignment and stop
            break
    return lab.astype(np.int32), it

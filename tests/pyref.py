"""Independent pure-Python restatement of src/em.rs (m_step :87-133, do_em :144-255),
written from the reference text, used to cross-pin the C oracle on small cases.
TEST INFRASTRUCTURE ONLY."""
MIN_READ_THRESH = 1e-5
EM_DENOM_THRESH = 1e-30


def m_step(rows, prev, curr, model_coverage=False):
    for txps, probs, covs in rows:
        denom = 0.0
        for t, p, cp in zip(txps, probs, covs):
            cov = cp if model_coverage else 1.0
            denom += prev[t] * float(p) * cov * 1.0
        if denom > EM_DENOM_THRESH:
            for t, p, cp in zip(txps, probs, covs):
                cov = cp if model_coverage else 1.0
                curr[t] += (prev[t] * float(p) * cov * 1.0) / denom


def do_em(make_rows, n_reads, n_txps, max_iter, thresh, min_iter=50, init=None, model_coverage=False):
    prev = list(init) if init is not None else [n_reads / n_txps] * n_txps
    curr = [0.0] * n_txps
    rel_diff = 0.0
    niter = 0
    while niter < max_iter:
        m_step(make_rows(), prev, curr, model_coverage)
        for i in range(n_txps):
            if prev[i] > MIN_READ_THRESH:
                rd = (curr[i] - prev[i]) / prev[i]
                rel_diff = max(rel_diff, rd)
        prev, curr = curr, prev
        for i in range(n_txps):
            curr[i] = 0.0
        if rel_diff < thresh and niter > min_iter:
            break
        niter += 1
        rel_diff = 0.0
    for i in range(n_txps):
        if prev[i] < MIN_READ_THRESH:
            prev[i] = 0.0
    m_step(make_rows(), prev, curr, model_coverage)
    return curr, niter


def rows_of(row_ptr, txp, prob, cov=None, inds=None):
    n = len(row_ptr) - 1
    order = range(n) if inds is None else inds
    for r in order:
        s, e = int(row_ptr[r]), int(row_ptr[r + 1])
        yield ([int(x) for x in txp[s:e]], [float(x) for x in prob[s:e]],
               [float(x) for x in cov[s:e]] if cov is not None else [0.0] * (e - s))

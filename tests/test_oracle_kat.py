"""CPU tests of the oracle (oracle/em_oracle.c): the reference ships no EM tests
(parity unpinned), so the restatement is pinned by an independent pure-Python
restatement, analytic known answers and the frozen golden fixtures."""
import os

import numpy as np
import pytest

from tests import pyref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def csr(rows):
    """rows: list of lists of (txp, prob)"""
    rp = np.zeros(len(rows) + 1, dtype=np.uint64)
    rp[1:] = np.cumsum([len(r) for r in rows])
    tx = np.array([t for r in rows for t, _ in r], dtype=np.uint32)
    pr = np.array([p for r in rows for _, p in r], dtype=np.float32)
    return rp, tx, pr


@pytest.mark.parametrize("min_iter", [50, 1])
def test_c_oracle_matches_pure_python_restatement(oracle_mod, tiny_store, min_iter):
    s = tiny_store
    c, niter, rel, sweeps = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, min_iter=min_iter)
    ref, ref_niter = pyref.do_em(lambda: pyref.rows_of(s.row_ptr, s.txp_id, s.prob), s.n_reads, s.n_txps, 1000, 1e-3,
                                 min_iter=min_iter)
    assert niter == ref_niter
    assert sweeps == niter + 1
    np.testing.assert_array_equal(c, np.array(ref))  # same order of f64 operations -> bit identical


def test_m_step_matches_python_with_coverage(oracle_mod, tiny_store):
    s = tiny_store
    rng = np.random.default_rng(0)
    cov = rng.uniform(0.1, 1.0, size=s.nnz)
    prev = rng.uniform(0.0, 5.0, size=s.n_txps)
    got = oracle_mod.m_step(s.row_ptr, s.txp_id, s.prob, prev, cov=cov)
    want = [0.0] * s.n_txps
    pyref.m_step(pyref.rows_of(s.row_ptr, s.txp_id, s.prob, cov), list(prev), want, model_coverage=True)
    np.testing.assert_array_equal(got, np.array(want))


def test_all_unique_reads_give_exact_read_counts(oracle_mod):
    rng = np.random.default_rng(1)
    t = rng.integers(0, 17, size=500)
    rows = [[(int(x), float(rng.uniform(0.1, 1.0)))] for x in t]
    rp, tx, pr = csr(rows)
    for max_iter in (1, 7, 1000):
        c, *_ = oracle_mod.do_em(rp, tx, pr, 17, max_iter=max_iter)
        np.testing.assert_array_equal(c, np.bincount(t, minlength=17).astype(np.float64))


def test_two_transcript_closed_form(oracle_mod):
    # n1 reads unique to A, n2 unique to B, n3 shared with equal probability:
    # fixed point a = n1*N/(n1+n2), b = n2*N/(n1+n2)
    n1, n2, n3 = 30, 10, 60
    rows = [[(0, 1.0)]] * n1 + [[(1, 1.0)]] * n2 + [[(0, 0.5), (1, 0.5)]] * n3
    rp, tx, pr = csr(rows)
    c, niter, rel, _ = oracle_mod.do_em(rp, tx, pr, 2, max_iter=100000, conv_thresh=1e-14)
    N = n1 + n2 + n3
    np.testing.assert_allclose(c, [n1 * N / (n1 + n2), n2 * N / (n1 + n2)], rtol=1e-9)


def test_counts_sum_to_assignable_reads(oracle_mod, small_store):
    s = small_store
    c, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    assert abs(c.sum() - s.n_reads) < 1e-6 * s.n_reads


def test_denominator_threshold_drops_read(oracle_mod):
    # prob 1e-37 * prev 1.0 <= 1e-30 -> the read contributes nothing (em.rs:115)
    rows = [[(0, 1.0)], [(1, 1e-37)], [(1, 1.0)]]
    rp, tx, pr = csr(rows)
    c, *_ = oracle_mod.do_em(rp, tx, pr, 3, max_iter=5)
    assert c.sum() == pytest.approx(2.0)


def test_stop_rule_minimum_iterations(oracle_mod):
    # a store that is converged from the first sweep: do_em stops at niter = 51
    # (52 loop sweeps), em_par's rule at niter = 2 (3 loop sweeps)  (em.rs:212 / :399)
    rows = [[(0, 1.0)]] * 10
    rp, tx, pr = csr(rows)
    _, niter, _, sweeps = oracle_mod.do_em(rp, tx, pr, 1, min_iter=50)
    assert (niter, sweeps) == (51, 52)
    _, niter, _, sweeps = oracle_mod.do_em(rp, tx, pr, 1, min_iter=1)
    assert (niter, sweeps) == (2, 3)


def test_max_iter_cap_and_zero(oracle_mod, tiny_store):
    s = tiny_store
    _, niter, _, sweeps = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, max_iter=7)
    assert (niter, sweeps) == (7, 7)
    c0, niter, _, sweeps = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, max_iter=0)
    assert (niter, sweeps) == (0, 0)
    # zero iterations = threshold the uniform start and sweep once
    want = oracle_mod.m_step(s.row_ptr, s.txp_id, s.prob, np.full(s.n_txps, s.n_reads / s.n_txps))
    np.testing.assert_array_equal(c0, want)


def test_signed_rel_diff_only_counts_increases(oracle_mod):
    # transcript 1 only ever loses mass to transcript 0: with the signed rule the
    # decrease of t1 is ignored and only t0's (shrinking) increase is tracked
    rows = [[(0, 1.0)]] * 50 + [[(0, 1.0), (1, 0.2)]] * 50
    rp, tx, pr = csr(rows)
    _, niter_signed, rel, _ = oracle_mod.do_em(rp, tx, pr, 2, min_iter=1, conv_thresh=1e-2)
    assert rel >= 0.0
    ref, ref_niter = pyref.do_em(lambda: pyref.rows_of(rp, tx, pr), 100, 2, 1000, 1e-2, min_iter=1)
    assert niter_signed == ref_niter


def test_init_zero_stays_zero(oracle_mod, tiny_store):
    s = tiny_store
    init = np.full(s.n_txps, s.n_reads / s.n_txps)
    init[::3] = 0.0
    c, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, init=init)
    assert np.all(c[::3] == 0.0)


def test_read_permutation_invariance(oracle_mod, tiny_store):
    s = tiny_store
    rng = np.random.default_rng(3)
    perm = rng.permutation(s.n_reads)
    lens = np.diff(s.row_ptr.astype(np.int64))
    rp2 = np.zeros_like(s.row_ptr); rp2[1:] = np.cumsum(lens[perm])
    idx = np.concatenate([np.arange(s.row_ptr[r], s.row_ptr[r + 1]) for r in perm]).astype(np.int64)
    c1, n1, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    c2, n2, *_ = oracle_mod.do_em(rp2, s.txp_id[idx], s.prob[idx], s.n_txps)
    assert n1 == n2
    np.testing.assert_allclose(c1, c2, rtol=1e-9, atol=1e-12)


def test_transcript_relabel_equivariance(oracle_mod, tiny_store):
    s = tiny_store
    rng = np.random.default_rng(4)
    relabel = rng.permutation(s.n_txps).astype(np.uint32)
    c1, n1, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    c2, n2, *_ = oracle_mod.do_em(s.row_ptr, relabel[s.txp_id], s.prob, s.n_txps)
    assert n1 == n2
    np.testing.assert_allclose(c2[relabel], c1, rtol=1e-12, atol=0)


def test_bootstrap_index_list_equals_weights(oracle_mod, tiny_store):
    s = tiny_store
    inds = oracle_mod.get_sample_inds(s.n_reads, 99)
    assert np.all(np.diff(inds.astype(np.int64)) >= 0) and inds.max() < s.n_reads and len(inds) == s.n_reads
    w = oracle_mod.inds_to_weights(inds, s.n_reads)
    assert w.sum() == s.n_reads
    c1, n1, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, inds=inds)
    c2, n2, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, wts=w)
    assert n1 == n2
    np.testing.assert_array_equal(c1, c2)  # sorted index list visits rows in the same order as weights


def test_bootstrap_identity_sample_equals_plain_em(oracle_mod, tiny_store):
    # the commented-out identity sampling of em.rs:278-283
    s = tiny_store
    inds = np.arange(s.n_reads, dtype=np.uint64)
    c1, n1, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, inds=inds)
    c2, n2, *_ = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    assert n1 == n2
    np.testing.assert_array_equal(c1, c2)


def test_threaded_port_matches_oracle(oracle_mod, small_store):
    s = small_store
    ps = oracle_mod.PortStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    c1, n1, _, sw1 = ps.em_par()
    c2, n2, _, sw2 = oracle_mod.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, min_iter=1)
    assert (n1, sw1) == (n2, sw2)
    np.testing.assert_allclose(c1, c2, rtol=1e-9, atol=1e-9)
    out, nit = ps.bootstrap(3, seed=5, nthreads=2)
    assert out.shape == (3, s.n_txps) and np.all(nit > 50)
    np.testing.assert_allclose(out.sum(axis=1), s.n_reads, rtol=1e-9)
    ps.close()


@pytest.mark.parametrize("name", ["sirv_store", "tiny_store"])
def test_oracle_reproduces_golden_fixture(oracle_mod, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    rp, tx, pr, M = g["row_ptr"], g["txp_id"], g["prob"], int(g["n_txps"])
    for mi in (50, 1):
        c, niter, rel, _ = oracle_mod.do_em(rp, tx, pr, M, min_iter=mi)
        assert niter == int(g[f"niter_min{mi}"])
        np.testing.assert_array_equal(c, g[f"counts_min{mi}"])
    for b in range(2):
        c, niter, *_ = oracle_mod.do_em(rp, tx, pr, M, wts=g[f"boot{b}_weights"])
        assert niter == int(g[f"boot{b}_niter"])
        np.testing.assert_array_equal(c, g[f"boot{b}_counts"])
    c, niter, *_ = oracle_mod.do_em(rp, tx, pr, M, cov=g["cov"])
    assert niter == int(g["niter_cov"])
    np.testing.assert_array_equal(c, g["counts_cov"])


def test_sirv_fixture_conserves_reads_per_gene(oracle_mod):
    # simulated reads only align inside their own SIRV gene, so the EM must hand every
    # gene exactly its reads back (isoform-level identifiability is a property of the
    # simulation, not of the EM, and is not asserted)
    g = np.load(os.path.join(GOLD, "sirv_store.npz"))
    genes = np.array([n[:5] for n in g["names"]])
    c = g["counts_min50"]
    true = np.bincount(g["true_txp"], minlength=int(g["n_txps"]))
    for gene in np.unique(genes):
        m = genes == gene
        assert c[m].sum() == pytest.approx(true[m].sum(), rel=1e-9)


def test_posteriors_and_aux_counts_known_answers(oracle_mod):
    rows = [[(0, 1.0)], [(0, 0.5), (1, 0.5)], [(1, 1.0), (2, 0.001)], [(2, 1.0)]]
    rp, tx, pr = csr(rows)
    counts = np.array([3.0, 1.0, 1.0])
    out, kept = oracle_mod.posteriors(rp, tx, pr, counts, 0.0)
    np.testing.assert_allclose(out, [1.0, 0.75, 0.25, 1.0 / 1.001, 0.001 / 1.001, 1.0], rtol=1e-7)  # probs are f32
    assert list(kept) == [1, 2, 2, 1]
    out, kept = oracle_mod.posteriors(rp, tx, pr, counts, 0.01)   # drops the 0.000999 alignment, renormalises
    np.testing.assert_allclose(out, [1.0, 0.75, 0.25, 1.0, 0.0, 1.0], rtol=1e-12)
    assert list(kept) == [1, 2, 1, 1]
    u, t = oracle_mod.aux_counts(rp, tx, 3)
    assert list(u) == [1, 0, 1] and list(t) == [2, 2, 2]


def test_coverage_model_known_answers(oracle_mod):
    """bulk.rs:103-108: add_interval -> logistic_prob -> normalize_read_probs on hand-computable inputs."""
    # one transcript of 300 nt, bin width 100, one read covering it entirely: all bins equal -> logistic(0) = 0.5,
    # a single-alignment read normalises to 1
    rp = np.array([0, 1], dtype=np.uint64)
    out = oracle_mod.coverage_model(rp, [0], [0], [300], [300], bin_width=100, growth_rate=2.0)
    np.testing.assert_allclose(out, [1.0])
    # two transcripts; read 0 hits both, reads 1..4 pile onto the first 100 nt of transcript 0 only
    rows_t = [0, 1, 0, 0, 0, 0]
    start = [0, 0, 0, 0, 0, 0]
    end = [300, 300, 150, 150, 150, 150]
    rp = np.array([0, 2, 3, 4, 5, 6], dtype=np.uint64)
    out = oracle_mod.coverage_model(rp, rows_t, start, end, [300, 300], bin_width=100, growth_rate=2.0)
    # transcript 0 bins: [0,100) gets 1 + 4*1 = 5, [100,200) gets 1 (end bin of the short reads is never visited),
    # [200,300) gets 1; total_weight 5 -> +0.05 each.  Transcript 1 bins: 1,1,1 (+0.01): flat -> 0.5 everywhere.
    c = np.array([5.05, 1.05, 1.05], dtype=np.float32).astype(np.float64)
    exp = c.sum() / 3
    p0 = np.clip(1 / (1 + np.exp(-2.0 * (exp - c) / exp)), 1e-8, 0.99999)
    e_t0_full = (p0[0] + p0[1]) / 2          # bins 0,1 visited (start_bin..end_bin excludes bin 2), weights 1,1
    e_t1_full = 0.5
    e_t0_short = p0[0]                        # start_bin 0, end_bin 1 -> only bin 0, w = 1
    want = [e_t0_full / (e_t0_full + e_t1_full), e_t1_full / (e_t0_full + e_t1_full), 1.0, 1.0, 1.0, 1.0]
    np.testing.assert_allclose(out, want, rtol=1e-12)
    assert e_t0_short > 0   # (single-alignment reads normalise to 1 whatever their coverage probability)


def test_statrs_ln_gamma_restated_matches_libm(oracle_mod):
    """statrs 0.18 ln_gamma (Lanczos g = 10.900511, 11 coefficients) as restated for the binomial coverage model:
    agrees with libm's lgamma on the argument range the model reaches (counts + 1, up to ~709 * bins)."""
    import math
    for x in [0.5, 1.0, 1.5, 2.0, 3.25, 10.0, 57.5, 710.0, 1234.567, 71000.0]:
        got, want = oracle_mod.statrs_ln_gamma(x), math.lgamma(x)
        assert abs(got - want) <= 5e-14 * max(1.0, abs(want)), (x, got, want)
    assert oracle_mod.statrs_ln_gamma(1.0) == pytest.approx(0.0, abs=1e-14)
    assert oracle_mod.statrs_ln_gamma(5.0) == pytest.approx(math.log(24.0), rel=1e-14)


def _binomial_probability_py(counts32, lengths32, rate):
    """binomial_probability (binomial_probability.rs:7-178) restated independently in numpy scalars."""
    import math
    f32 = np.float32
    count_sum = f32(0)
    for c in counts32:
        count_sum = f32(count_sum + c)
    if count_sum == 0 or rate == 0:
        return [0.0] * len(counts32)
    probs = [0.0 if (c == 0 or l == 0) else float(c) / (float(l) * rate) for c, l in zip(counts32, lengths32)]
    mx = max(counts32)
    mod = [f32(709.0) if c == mx else f32((float(c) * 709.0) / float(mx)) for c in counts32]
    sum_vec = f32(0)
    for m in mod:
        sum_vec = f32(sum_vec + m)
    ln1 = math.lgamma(float(sum_vec) + 1.0)
    res = []
    for p, m in zip(probs, mod):
        rest = f32(sum_vec - m)
        den = math.lgamma(float(m) + 1.0) + math.lgamma(float(rest) + 1.0)
        n2 = (math.log(p) if p > 1e-20 else math.log(1e-20)) * float(m)
        n3 = (math.log(1.0 - p) if (1.0 - p) > 1e-20 else math.log(1e-20)) * float(rest)
        res.append(math.exp(ln1 - den + n2 + n3))
    tot = sum(res)
    return [r / tot for r in res]


def test_binomial_coverage_model_known_answers(oracle_mod):
    """single_cell.rs:132-137: add_interval -> binomial_continuous_prob -> normalize_read_probs."""
    f32 = np.float32
    # flat coverage: every bin has the same count -> the same probability, 1/3 each; a single-alignment read -> 1
    rp = np.array([0, 1], dtype=np.uint64)
    out = oracle_mod.coverage_model_binomial(rp, [0], [0], [300], [300], bin_width=100)
    np.testing.assert_allclose(out, [1.0])
    # the two-transcript pile-up of the logistic test, bins of transcript 0 = 5.05, 1.05, 1.05 (f32)
    rows_t = [0, 1, 0, 0, 0, 0]
    start = [0, 0, 0, 0, 0, 0]
    end = [300, 300, 150, 150, 150, 150]
    rp = np.array([0, 2, 3, 4, 5, 6], dtype=np.uint64)
    out = oracle_mod.coverage_model_binomial(rp, rows_t, start, end, [300, 300], bin_width=100)
    c0 = [f32(5.05), f32(1.05), f32(1.05)]
    lens = [f32(100), f32(100), f32(100)]
    rate0 = sum(float(c) / float(l) for c, l in zip(c0, lens))
    p0 = _binomial_probability_py(c0, lens, rate0)
    c1 = [f32(1.01)] * 3
    p1 = _binomial_probability_py(c1, lens, sum(float(c) / 100.0 for c in c1))
    np.testing.assert_allclose(p1, [1 / 3] * 3, rtol=1e-12)
    assert abs(sum(p0) - 1.0) < 1e-12
    e_t0_full = (p0[0] + p0[1]) / 2           # bins 0 and 1 visited (start_bin..end_bin excludes the end bin)
    e_t1_full = (p1[0] + p1[1]) / 2
    want = [e_t0_full / (e_t0_full + e_t1_full), e_t1_full / (e_t0_full + e_t1_full), 1.0, 1.0, 1.0, 1.0]
    np.testing.assert_allclose(out, want, rtol=1e-10)
    # a transcript nobody aligns to has no counts: probabilities 0, and nothing divides by zero elsewhere
    out = oracle_mod.coverage_model_binomial(np.array([0, 1], dtype=np.uint64), [1], [0], [250], [300, 250], bin_width=100)
    np.testing.assert_allclose(out, [1.0])


def test_filter_known_answers(oracle_mod):
    """AlignmentFilters::filter (oarfish_types.rs:955-1130) on hand-made groups."""
    import math
    f32 = np.float32
    txp_len = [1000, 1000, 1000]
    # group 0: best score 100 on txp 0; 96 kept (>= 0.95 * 100), 90 dropped by the score threshold; one supplementary,
    #          one unmapped, one too short.  group 1: only an unmapped record.  group 2: best score 0 -> no valid alignment.
    # group 3: aligned fraction 100/1000 < 0.5 -> dropped.  group 4: reverse strand only, forward-only filter.
    gp = [0, 6, 7, 8, 9, 10]
    ref = [0, 1, 2, 1, 0, 2, 0, 1, 2, 0]
    start = [1, 1, 1, 1, 1, 1, 1, 1, 1, 1]
    span = [500, 500, 500, 500, 500, 10, 500, 500, 100, 500]
    end = [s0 + sp - 1 for s0, sp in zip(start, span)]
    score = [100, 96, 90, 99, 99, 99, 50, 0, 80, 70]
    flags = [0, 0, 0, 4, 1, 0, 1, 0, 0, 2]
    seq = [600, 0, 0, 0, 0, 0, 600, 600, 1000, 600]
    rp, tx, pr, src, grp, disc = oracle_mod.filter_records(gp, ref, start, end, span, score, flags, seq, txp_len, which_strand=1,
                                                           min_aligned_len=50, min_aligned_fraction=0.5, score_threshold=0.95)
    assert list(rp) == [0, 2] and list(tx) == [0, 1] and list(src) == [0, 1] and list(grp) == [0]
    want = [f32(1.0), f32(math.exp(float(f32(f32(96 - 100) / f32(5.0)))))]
    np.testing.assert_allclose(pr, want, rtol=1e-6)
    assert disc == {"discard_5p": 0, "discard_3p": 0, "discard_score": 1, "discard_aln_frac": 1, "discard_aln_len": 1,
                    "discard_ori": 1, "discard_supp": 1, "no_mapping": 1, "no_valid_aln": 2, "valid_best_aln": 1}
    # clipping: with three_prime_clip = 100 an alignment must end beyond len - 100; with five_prime_clip = 50 it must start before 50
    rp, tx, pr, src, grp, disc = oracle_mod.filter_records([0, 3], [0, 0, 0], [1, 60, 1], [950, 990, 800], [950, 931, 800], [10, 10, 10],
                                                           [0, 0, 0], [1000, 0, 0], [1000], three_prime_clip=100, five_prime_clip=50)
    assert list(src) == [0] and disc["discard_5p"] == 1 and disc["discard_3p"] == 1

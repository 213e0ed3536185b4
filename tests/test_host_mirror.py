"""Host-side mirror of the reference interface (oarfish_b200/em.py): layout and
flattening logic, no GPU needed."""
import numpy as np

from oarfish_b200 import ALN_INFO_DTYPE, AlignmentFilters, InMemoryAlignmentStore


def test_aln_info_is_24_bytes_like_the_rust_struct():
    assert ALN_INFO_DTYPE.itemsize == 24  # oarfish_types.rs:330-337


def make_group(ids, probs):
    a = np.zeros(len(ids), dtype=ALN_INFO_DTYPE)
    a["ref_id"] = ids
    a["start"] = 10
    a["end"] = 500
    return a, np.array(probs, dtype=np.float32)


def test_add_filtered_group_builds_boundaries_like_the_reference():
    s = InMemoryAlignmentStore()
    assert s.len() == 0 and s.total_len() == 0
    assert s.add_filtered_group(*make_group([3, 1], [1.0, 0.5]))
    assert not s.add_filtered_group(*make_group([], []))  # empty groups are dropped (oarfish_types.rs:724)
    assert s.add_filtered_group(*make_group([2], [1.0]))
    assert s.len() == 2 and s.num_aligned_reads() == 2 and s.total_len() == 3
    rp, txp, prob, aux = s.csr()
    np.testing.assert_array_equal(rp, [0, 2, 3])
    np.testing.assert_array_equal(txp, [3, 1, 2])
    np.testing.assert_array_equal(prob, np.array([1.0, 0.5, 1.0], dtype=np.float32))
    assert aux is None  # model_coverage off -> factor 1.0 (em.rs:108)
    assert rp.dtype == np.uint64 and txp.dtype == np.uint32 and prob.dtype == np.float32
    rows = list(s.iter())
    assert [len(r[0]) for r in rows] == [2, 1]
    assert np.all(s.coverage_probabilities == 0.0)  # zeros until the coverage model runs (:731)


def test_random_sampling_iter_repeats_rows():
    s = InMemoryAlignmentStore()
    s.add_filtered_group(*make_group([0], [1.0]))
    s.add_filtered_group(*make_group([1, 2], [1.0, 0.3]))
    rows = list(s.random_sampling_iter([1, 1, 0]))
    assert [list(r[0]["ref_id"]) for r in rows] == [[1, 2], [1, 2], [0]]


def test_model_coverage_exports_aux():
    rp = np.array([0, 2, 3], dtype=np.uint64)
    s = InMemoryAlignmentStore.from_csr(rp, np.array([0, 1, 1], np.uint32), np.ones(3, np.float32),
                                        coverage=np.array([0.5, 0.25, 1.0]), model_coverage=True)
    _, _, _, aux = s.csr()
    np.testing.assert_array_equal(aux, [0.5, 0.25, 1.0])
    s2 = InMemoryAlignmentStore(AlignmentFilters(model_coverage=False))
    assert s2.csr()[3] is None


def test_cpp_host_mirror_selfcheck():
    """include/oarfish_em.hpp: the C++ mirror of em::em / em_par / bootstrap and of the store types."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oarfish_b200", "lib", "host_mirror_test")
    assert os.path.exists(exe), "run `make product`"
    res = subprocess.run([exe, "--selfcheck"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and "selfcheck ok" in res.stdout


def test_oarstore_roundtrip(tmp_path):
    from oarfish_b200 import storefile, synth
    s = synth.make_config("tiny")
    aux = np.linspace(0.1, 1.0, s.nnz)
    for a in (None, aux):
        p = str(tmp_path / ("a.oarstore" if a is None else "b.oarstore"))
        storefile.write_store(p, s.row_ptr, s.txp_id, s.prob, s.n_txps, aux=a)
        rp, tx, pr, m, ax = storefile.read_store(p)
        assert m == s.n_txps and np.array_equal(rp, s.row_ptr) and np.array_equal(tx, s.txp_id) and np.array_equal(pr, s.prob)
        assert (ax is None) == (a is None) and (a is None or np.array_equal(ax, a))
    with open(tmp_path / "bad", "wb") as f:
        f.write(b"x" * 100)
    import pytest
    with pytest.raises(ValueError):
        storefile.read_store(str(tmp_path / "bad"))


def test_oarstore_carries_the_em_answer(tmp_path):
    """A dump of the reference's own run (the Rust side writes store + counts + EM parameters)."""
    from oarfish_b200 import storefile, synth
    s = synth.make_config("tiny")
    counts = np.linspace(0.0, 5.0, s.n_txps)
    p = str(tmp_path / "ref.oarstore")
    storefile.write_store(p, s.row_ptr, s.txp_id, s.prob, s.n_txps, counts=counts, min_iter=50, max_iter=1000, conv_thresh=1e-3)
    rp, tx, pr, m, ax, ref = storefile.read_store_full(p)
    assert ax is None and m == s.n_txps
    np.testing.assert_array_equal(rp, s.row_ptr)
    np.testing.assert_array_equal(ref["counts"], counts)
    assert (ref["min_iter"], ref["max_iter"], ref["conv_thresh"], ref["niter"], ref["from_oracle"]) == (50, 1000, 1e-3, None, False)
    assert storefile.read_store(p)[3] == s.n_txps            # the five-tuple reader ignores the answer
    assert storefile.read_store_full(p.replace("ref", "ref"))[5] is not None


def test_golden_oarstore_fixture_matches_the_oracle(oracle_mod=None):
    """tests/golden/sirv_shaped.oarstore: store + the ORACLE's answer (flag bit 2), written by tests/golden/make_golden.py.
    It keeps the dump reader and the GPU test that consumes dumps exercised until a maintainer adds a file written
    by the Rust binary (flag bit 2 clear)."""
    import glob
    import os
    from oarfish_b200 import storefile
    from oracle import oracle
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.oarstore")))
    assert files, "tests/golden holds no .oarstore fixture"
    for f in files:
        rp, tx, pr, m, ax, ref = storefile.read_store_full(f)
        assert ref is not None
        want, niter, _, _ = oracle.do_em(np.asarray(rp), np.asarray(tx), np.asarray(pr), m, max_iter=ref["max_iter"],
                                         conv_thresh=ref["conv_thresh"], min_iter=ref["min_iter"],
                                         cov=None if ax is None else np.asarray(ax))
        rtol = 1e-12 if ref["from_oracle"] else 1e-5
        big = want > 1e-8
        assert (np.abs(np.asarray(ref["counts"])[big] - want[big]) / want[big]).max() <= rtol
        if ref["niter"] is not None:
            assert ref["niter"] == niter


def test_synthetic_workload_generators():
    """The generators behind bench.py / tools (no GPU): shuffled ids, long rows, sparse cells, raw alignment records."""
    from oarfish_b200 import synth
    s = synth.make_config("tiny")
    p = synth.permute_ids(s, 3)
    assert p.nnz == s.nnz and np.array_equal(np.sort(np.unique(p.txp_id)), np.sort(np.unique(p.txp_id)))
    assert np.array_equal(p.row_ptr, s.row_ptr) and not np.array_equal(p.txp_id, s.txp_id)
    l = synth.with_long_rows(s, 0.1, 128, 200, 4)
    lens, lens0 = np.diff(l.row_ptr.astype(np.int64)), np.diff(s.row_ptr.astype(np.int64))
    assert lens.max() >= 128 and (lens != lens0).mean() > 0.05 and int(l.row_ptr[-1]) == l.nnz == len(l.txp_id) == len(l.prob)
    rp = l.row_ptr.astype(np.int64)
    i = int(np.argmax(lens))
    assert len(np.unique(l.txp_id[rp[i]:rp[i + 1]])) == lens[i] and l.txp_id.max() < l.n_txps   # distinct transcripts within a read
    st, crp = synth.make_cells([300, 0, 500], 4000, 5.0, 9, expressed=400)
    assert list(crp) == [0, 300, 300, 800] and st.n_reads == 800
    a0, a1 = int(st.row_ptr[0]), int(st.row_ptr[300])
    assert len(np.unique(st.txp_id[a0:a1])) <= 400 and st.txp_id.max() < 4000
    rec = synth.make_records(2000, 300, 1)
    assert int(rec["group_ptr"][-1]) == len(rec["score"]) == len(rec["flags"]) == len(rec["ref_id"])
    assert rec["ref_id"].max() < 300 and np.all(rec["aln_end"] >= rec["aln_start"])
    assert np.all(rec["aln_end"].astype(np.int64) <= rec["txp_len"][rec["ref_id"]].astype(np.int64) + 0)

#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run once, in the build container).

  sirv_store.npz   config C1 of BASELINE.md: a SIRV-*shaped*, SIMULATED store.  The
                   reference ships no reads/BAM and no aligner is installed, so the
                   store is built from the reference's test DATA files only
                   (test_data/SIRV_isoforms_multi-fasta-annotation_C_170612a.gtf for the
                   69 isoform exon chains, test_data/molar_concentrations.xlsx sheet3
                   for the E2 mix molarities): 10 000 simulated long reads, each a
                   random sub-interval of its true isoform, "aligned" to every isoform
                   of the same gene whose exons cover >= 95 % of the read, with
                   prob = expf((score - best)/5) as in oarfish_types.rs:1107-1118.
  tiny_store.npz   the synthetic "tiny" store (oarfish_b200.synth) for regression.
  sirv_shaped.oarstore  the SIRV-shaped store as an .oarstore dump carrying the EM's answer (oarfish_b200/storefile.py)

Each fixture holds the inputs AND the oracle's outputs (counts / niter for both
stop rules, two index-list bootstrap replicates), so the tests need neither
/root/reference nor a re-run of this script.  The oracle is the CPU restatement
in oracle/ (parity unpinned by the reference: it has no EM tests).
"""
import os
import re
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/test_data"

from oracle import oracle  # noqa: E402
from oarfish_b200 import synth  # noqa: E402


def read_gtf(path):
    txps = {}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if len(f) < 9 or f[2] != "exon":
            continue
        gene = re.search(r'gene_id "([^"]+)"', f[8]).group(1)
        tid = re.search(r'transcript_id "([^"]+)"', f[8]).group(1)
        txps.setdefault(tid, {"gene": gene, "strand": f[6], "exons": []})["exons"].append((int(f[3]), int(f[4]) + 1))
    for t in txps.values():
        t["exons"].sort()
        t["len"] = sum(e - s for s, e in t["exons"])
    return txps


def read_molarity(path):
    z = zipfile.ZipFile(path)
    strings = [re.sub(r"<[^>]+>", "", x) for x in re.findall(r"<si>(.*?)</si>", z.read("xl/sharedStrings.xml").decode(), re.S)]
    sheet = z.read("xl/worksheets/sheet3.xml").decode()
    out = {}
    for row in re.findall(r"<row [^>]*>(.*?)</row>", sheet, re.S):
        vals = []
        for attr, body in re.findall(r"<c r=\"[A-Z]+\d+\"([^>]*?)(?:/>|>(.*?)</c>)", row, re.S):
            v = re.search(r"<v>(.*?)</v>", body or "")
            val = v.group(1) if v else None
            if 't="s"' in attr and val is not None:
                val = strings[int(val)]
            vals.append(val)
        if len(vals) >= 2 and vals[0] and vals[0].startswith("SIRV"):
            out[vals[0]] = float(vals[1])
    return out


def tx_interval_to_blocks(exons, a, b):
    """transcript-coordinate interval [a,b) -> genomic blocks"""
    blocks, off = [], 0
    for s, e in exons:
        l = e - s
        lo, hi = max(a, off), min(b, off + l)
        if lo < hi:
            blocks.append((s + lo - off, s + hi - off))
        off += l
    return blocks


def covered(blocks, exons):
    c = 0
    for bs, be in blocks:
        for s, e in exons:
            lo, hi = max(bs, s), min(be, e)
            if lo < hi:
                c += hi - lo
    return c


def make_sirv(n_reads=10_000, seed=1):
    txps = read_gtf(os.path.join(REF, "SIRV_isoforms_multi-fasta-annotation_C_170612a.gtf"))
    mol = read_molarity(os.path.join(REF, "molar_concentrations.xlsx"))
    names = sorted(txps)
    by_gene = {}
    for i, n in enumerate(names):
        by_gene.setdefault(txps[n]["gene"], []).append(i)
    ab = np.array([mol.get(n, 0.0) for n in names])
    ab = ab / ab.sum()
    rng = np.random.default_rng(seed)
    true_t = rng.choice(len(names), size=n_reads, p=ab)
    row_ptr, txp, prob, start, end = [0], [], [], [], []
    for t in true_t:
        T = txps[names[t]]
        L = T["len"]
        rl = int(rng.uniform(0.4, 1.0) * L)
        a = int(rng.integers(0, L - rl + 1))
        blocks = tx_interval_to_blocks(T["exons"], a, a + rl)
        hits = []
        for j in by_gene[T["gene"]]:
            if txps[names[j]]["strand"] != T["strand"]:
                continue
            c = covered(blocks, txps[names[j]]["exons"])
            if c >= 0.95 * rl:
                hits.append((j, c))
        best = max(c for _, c in hits)
        order = rng.permutation(len(hits))
        for k in order:
            j, c = hits[k]
            txp.append(j)
            prob.append(np.exp(np.float32((c - best) / 5.0), dtype=np.float32))
            start.append(a); end.append(a + rl)
        row_ptr.append(len(txp))
    return dict(row_ptr=np.array(row_ptr, dtype=np.uint64), txp_id=np.array(txp, dtype=np.uint32),
                prob=np.array(prob, dtype=np.float32), n_txps=np.int64(len(names)), names=np.array(names),
                true_txp=true_t.astype(np.uint32), molarity=ab)


def add_oracle_outputs(d):
    rp, tx, pr, M = d["row_ptr"], d["txp_id"], d["prob"], int(d["n_txps"])
    for mi in (50, 1):
        c, niter, rel, sweeps = oracle.do_em(rp, tx, pr, M, max_iter=1000, conv_thresh=1e-3, min_iter=mi)
        d[f"counts_min{mi}"] = c
        d[f"niter_min{mi}"] = np.int64(niter)
        d[f"rel_min{mi}"] = np.float64(rel)
    n = len(rp) - 1
    for b in range(2):
        inds = oracle.get_sample_inds(n, 1000 + b)
        c, niter, _, _ = oracle.do_em(rp, tx, pr, M, max_iter=1000, conv_thresh=1e-3, min_iter=50, inds=inds)
        d[f"boot{b}_weights"] = oracle.inds_to_weights(inds, n)
        d[f"boot{b}_counts"] = c
        d[f"boot{b}_niter"] = np.int64(niter)
    # coverage-model variant: a deterministic per-alignment f64 factor
    cov = (0.25 + 0.75 * ((np.arange(len(tx)) * 2654435761 % 1000) / 999.0)).astype(np.float64)
    c, niter, _, _ = oracle.do_em(rp, tx, pr, M, max_iter=1000, conv_thresh=1e-3, min_iter=50, cov=cov)
    d["cov"] = cov
    d["counts_cov"] = c
    d["niter_cov"] = np.int64(niter)
    return d


def main():
    sirv = add_oracle_outputs(make_sirv())
    np.savez_compressed(os.path.join(HERE, "sirv_store.npz"), **sirv)
    print("sirv:", len(sirv["row_ptr"]) - 1, "reads", len(sirv["txp_id"]), "alignments", int(sirv["n_txps"]), "txps",
          "niter", int(sirv["niter_min50"]), int(sirv["niter_min1"]))
    # the same store as an .oarstore dump WITH the EM's answer -- here the oracle's (flag bit 2 set); a file written by
    # the Rust binary has the same layout with the bit clear and pins parity on the reference itself
    from oarfish_b200 import storefile
    storefile.write_store(os.path.join(HERE, "sirv_shaped.oarstore"), sirv["row_ptr"], sirv["txp_id"], sirv["prob"], int(sirv["n_txps"]),
                          counts=sirv["counts_min50"], min_iter=50, max_iter=1000, conv_thresh=1e-3, niter=int(sirv["niter_min50"]),
                          from_oracle=True)
    s = synth.make_config("tiny")
    tiny = add_oracle_outputs(dict(row_ptr=s.row_ptr, txp_id=s.txp_id, prob=s.prob, n_txps=np.int64(s.n_txps)))
    np.savez_compressed(os.path.join(HERE, "tiny_store.npz"), **tiny)
    print("tiny:", s.n_reads, "reads", s.nnz, "alignments niter", int(tiny["niter_min50"]), int(tiny["niter_min1"]))


if __name__ == "__main__":
    main()

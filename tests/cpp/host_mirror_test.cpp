// Drives the C++ host mirror (include/oarfish_em.hpp) the way oarfish's bulk driver drives src/em.rs
// (bulk.rs:131-193).  Usage: host_mirror_test <store.bin> <out.bin> [seed]
//   store.bin: u64 n_reads, u64 nnz, u64 n_txps, u64 row_ptr[n_reads+1], u32 txp[nnz], f32 prob[nnz]
//   out.bin  : f64 em[M], f64 em_par[M], f64 boot[2][M], f64 boot_multi[2][M]
//              (boot_multi: the same two replicates through oar_multi_bootstrap on all visible devices, from two
//               concurrent host threads with their own stores -- distinct handles are independent)
// With "--selfcheck" only the host logic runs (no GPU needed).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <cmath>
#include <thread>

#include "oarfish_em.hpp"

using namespace oarfish;

static int selfcheck()
{
    InMemoryAlignmentStore st;
    if (st.len() != 0 || st.total_len() != 0) return 1;
    AlnInfo a; a.ref_id = 3; a.start = 10; a.end = 510;
    AlnInfo b; b.ref_id = 1; b.start = 0; b.end = 100;
    if (!st.add_filtered_group({a, b}, {1.0f, 0.5f})) return 2;
    if (st.add_filtered_group({}, {})) return 3;                 // empty groups are dropped (oarfish_types.rs:724)
    if (!st.add_filtered_group({b}, {1.0f})) return 4;
    if (st.len() != 2 || st.num_aligned_reads() != 2 || st.total_len() != 3) return 5;
    if (st.boundaries() != std::vector<size_t>({0, 2, 3})) return 6;
    if (a.alignment_span() != 500) return 7;
    if (st.coverage_probabilities != std::vector<double>({0.0, 0.0, 0.0})) return 8;
    std::puts("selfcheck ok");
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "--selfcheck")) return selfcheck();
    if (argc < 3) { std::fprintf(stderr, "usage: %s <store.bin> <out.bin> [seed]\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    uint64_t hdr[3];
    in.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    const uint64_t n_reads = hdr[0], nnz = hdr[1], n_txps = hdr[2];
    std::vector<uint64_t> row_ptr(n_reads + 1);
    std::vector<uint32_t> txp(nnz);
    std::vector<float> prob(nnz);
    in.read(reinterpret_cast<char *>(row_ptr.data()), sizeof(uint64_t) * row_ptr.size());
    in.read(reinterpret_cast<char *>(txp.data()), sizeof(uint32_t) * nnz);
    in.read(reinterpret_cast<char *>(prob.data()), sizeof(float) * nnz);
    if (!in) { std::fprintf(stderr, "short read\n"); return 2; }

    // build the store group by group, as parse_alignments -> add_group does (alignment_parser.rs:301-437)
    InMemoryAlignmentStore store;
    for (uint64_t r = 0; r < n_reads; ++r) {
        std::vector<AlnInfo> alns;
        std::vector<float> ps;
        for (uint64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) {
            AlnInfo a; a.ref_id = txp[j]; a.start = 0; a.end = 1;
            alns.push_back(a); ps.push_back(prob[j]);
        }
        store.add_filtered_group(alns, ps);
    }
    std::vector<TranscriptInfo> txps(n_txps);
    EMInfo emi;
    emi.eq_map = &store; emi.txp_info = &txps; emi.max_iter = 1000; emi.convergence_thresh = 1e-3;
    try {
        const std::vector<double> c1 = em(emi, 1);          // bulk.rs:157-158
        const std::vector<double> c2 = em_par(emi, 8);      // bulk.rs:155-156
        const uint64_t seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 7;
        const auto reps = bootstrap(emi, 2, 4, seed);       // bulk.rs:179
        std::ofstream out(argv[2], std::ios::binary);
        out.write(reinterpret_cast<const char *>(c1.data()), sizeof(double) * c1.size());
        out.write(reinterpret_cast<const char *>(c2.data()), sizeof(double) * c2.size());
        for (const auto &r : reps) out.write(reinterpret_cast<const char *>(r.data()), sizeof(double) * r.size());
        // em::bootstrap over every visible device, called from two host threads at once (each with its own copy of
        // the store, like the workers of single_cell.rs:91-193): both must reproduce the single-device replicates
        const int ndev = oar_device_count();
        std::vector<int> devs;
        for (int d = 0; d < (ndev > 0 ? ndev : 1); ++d) devs.push_back(d);
        std::vector<std::vector<double>> got[2];
        std::string err[2];
        auto worker = [&](int k) {
            try {
                InMemoryAlignmentStore mine = store;   // deep copy, own device handles
                EMInfo e2 = emi; e2.eq_map = &mine; e2.devices = devs;
                if (devs.size() == 1) { em(e2, 1); }
                got[k] = bootstrap(e2, 2, 4, seed);
            } catch (const std::exception &e) { err[k] = e.what(); }
        };
        std::thread t0(worker, 0), t1(worker, 1);
        t0.join(); t1.join();
        for (int k = 0; k < 2; ++k) if (!err[k].empty()) throw std::runtime_error("worker: " + err[k]);
        for (size_t b = 0; b < 2; ++b)
            for (size_t i = 0; i < got[0][b].size(); ++i)
                if (got[0][b][i] != got[1][b][i] && std::abs(got[0][b][i] - got[1][b][i]) > 1e-9 * std::abs(got[0][b][i]))
                    throw std::runtime_error("concurrent bootstraps disagree");
        for (const auto &r : got[0]) out.write(reinterpret_cast<const char *>(r.data()), sizeof(double) * r.size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
    return 0;
}

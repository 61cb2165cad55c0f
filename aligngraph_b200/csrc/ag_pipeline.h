// One unit (chromosome or --part slice) through the hot path: the five calls of the reference's loop body (AG:4768-4776)
// expressed over an Engine.  The product instantiates it with AgDevice (CUDA, ag_device.cu); the CPU test-suite instantiates
// it with a host emulation of the same per-thread code (tests/emul) to check the formulation without a GPU.
#pragma once
#include "ag_host.h"
#include "ag_device.cuh"
#include <chrono>
#include <cstdlib>
#include <cstdio>

struct AgUnitResult {
    std::string initial_text; AgText pre_text, ext_text;  // tmp/_initial_contigs.N.fa, _pre_extended_contigs.N.fa, _extended_contigs.N.fa
    double t_parse = 0, t_device = 0, t_post = 0;  // wall seconds: host parsing / device section incl. copies / host post passes
    u64 n_aln = 0, n_walks = 0, n_emitted = 0;
    void reset() { initial_text.clear(); pre_text.clear(); ext_text.clear(); t_parse = t_device = t_post = 0; n_aln = n_walks = n_emitted = 0; }   // keeps the buffers' capacity
};

inline AgUnitInput ag_unit_input(const AgUnit& u) {
    AgUnitInput in;
    in.ref = u.ref.data(); in.n_ref = u.n_ref; in.n_pos = (u32)u.ref.size();
    in.threads = u.threads.empty() ? nullptr : u.threads.data(); in.n_threads = (u32)u.threads.size();
    const bool table = u.cm_start.size() == u.ref.size() + 1 && u.cm.size() == u.chain_pos.size();   // explicit table (ag_set_contimers / expanded)
    in.cm_start = table ? u.cm_start.data() : nullptr; in.cm = table ? u.cm.data() : nullptr; in.n_cm = (u32)u.chain_pos.size();
    in.chain_pos = u.chain_pos.data(); in.chain_base = u.chain_base.data();
    in.aln = u.aln.data(); in.n_aln = u.aln.size(); in.ext = u.ext.data(); in.n_ext = u.ext.size();
    if (!u.cdesc.empty() && u.chain_pos.empty()) {   // run-space contig threads: the device expands the chain-major arrays itself
        in.cdesc = u.cdesc.data(); in.n_desc = (u32)u.cdesc.size(); in.cruns = u.cruns.data(); in.n_runs = (u32)u.cruns.size();
        in.contig_blob = u.chunks->blob.data(); in.blob_bytes = u.chunks->blob.size(); in.blob_version = u.chunks->version;
        in.n_cm = u.n_cm_runs; in.chain_pos = nullptr; in.chain_base = nullptr;
    }
    in.aln_on_device = u.aln_on_device;
    if (u.aln_on_device) { in.aln = nullptr; in.n_aln = 0; in.ext = nullptr; in.n_ext = 0; }
    return in;
}

// loadGenome + loadContigAlignment + the parsing half of loadReadAlignment (AG:4768-4772).  parse_sam == false leaves the SAM to the caller
// (the product parses it on the device, AgDevice::ingest_sam, and only falls back to ag_parse_sam for files outside the well-formed layout)
// host_chains == false leaves the chain-major contiMer arrays to the device as well (k_chain_expand over the run-space contig threads)
inline void ag_prepare_unit(const AgReads& reads, const std::string& tmp, int unit, AgUnit& u, std::string& initial_text, bool parse_sam = true, bool host_chains = true) {
    std::string n = std::to_string(unit);
    auto t0 = std::chrono::steady_clock::now();
    ag_load_genome(tmp + "/_genome." + n + ".fa", u);
    auto t1 = std::chrono::steady_clock::now();
    // contig threads in run space (interval arithmetic on the PSL blocks); the per-base formulation takes inputs the run form does not model
    if (getenv("AG_CONTIG_PERBASE") || !ag_thread_contigs_runs(tmp + "/_contigs.fa", tmp + "/_contigs_genome." + n + ".psl", initial_text, u)) {
        u.cdesc.clear(); u.cruns.clear(); u.chunks.reset(); u.n_cm_runs = 0;
        ag_thread_contigs(tmp + "/_contigs.fa", tmp + "/_contigs_genome." + n + ".psl", initial_text, u);
    } else if (host_chains) ag_expand_chains(u);
    auto t2 = std::chrono::steady_clock::now();
    if (parse_sam) ag_parse_sam(tmp + "/_reads_genome." + n + ".bowtie", reads, u);
    auto t3 = std::chrono::steady_clock::now();
    if (getenv("AG_POST_TIMING")) fprintf(stderr, "[parse] genome %.1f ms, contigs %.1f ms, sam %.1f ms\n", std::chrono::duration<double>(t1 - t0).count() * 1e3,
                                          std::chrono::duration<double>(t2 - t1).count() * 1e3, std::chrono::duration<double>(t3 - t2).count() * 1e3);
}

// extendContigs + scaffoldContigs (AG:4774-4776) on a built graph.  The device materialises and copies the contig bases while the host
// already prepares everything that does not read them (contig records, header lines, text offsets).
template <class Engine> void ag_extend_unit(Engine& eng, const AgReads& reads, const std::string& ref, AgUnitResult& r, u64& n_walks, u64& n_emitted) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<ag_walk> walks;
    eng.extend(walks);
    std::vector<u32> sel;
    ag_select_emitted(walks, sel);
    char* bases = nullptr; std::vector<u64> offs;   // bases: engine-owned buffer (page-locked on the device), valid after materialize_wait until the next materialize
    eng.materialize_begin(walks, sel, bases, offs);
    eng.occupancy_begin();
    std::vector<AgContig> contigs; AgMakeState ms;
    ag_make_contigs_begin(walks, sel, bases, offs, contigs, ms);
    eng.materialize_wait();
    auto t1 = std::chrono::steady_clock::now();
    ag_make_contigs_finish(walks, sel, bases, offs, reads, contigs, ms, r.pre_text);
    auto t1a = std::chrono::steady_clock::now();
    ag_dedup_join(contigs);
    auto t1b = std::chrono::steady_clock::now();
    std::vector<unsigned char> occ;
    eng.occupancy_wait(occ);
    ag_scaffold(contigs, ref, occ, r.ext_text);
    auto t2 = std::chrono::steady_clock::now();
    if (getenv("AG_POST_TIMING")) fprintf(stderr, "[post] make (second half) %.2f ms, dedup_join %.2f ms, scaffold %.2f ms\n", std::chrono::duration<double>(t1a - t1).count() * 1e3,
                                          std::chrono::duration<double>(t1b - t1a).count() * 1e3, std::chrono::duration<double>(t2 - t1b).count() * 1e3);
    r.t_device += std::chrono::duration<double>(t1 - t0).count();
    r.t_post += std::chrono::duration<double>(t2 - t1).count();
    n_walks = walks.size(); n_emitted = sel.size();
}

// graph build + extendContigs + scaffoldContigs on prepared arrays
template <class Engine> void ag_process_unit(Engine& eng, const AgReads& reads, const AgUnit& u, AgUnitResult& r) {
    auto t0 = std::chrono::steady_clock::now();
    eng.load_unit(ag_unit_input(u));
    eng.build();
    r.t_device += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ag_extend_unit(eng, reads, u.ref, r, r.n_walks, r.n_emitted);
    r.n_aln = u.aln.size();
}

template <class Engine> void ag_run_unit_files(Engine& eng, const AgReads& reads, const std::string& tmp, int unit, AgUnitResult& r, bool write = true) {
    auto t0 = std::chrono::steady_clock::now();
    AgUnit u;
    ag_prepare_unit(reads, tmp, unit, u, r.initial_text);
    r.t_parse += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ag_process_unit(eng, reads, u, r);
    if (write) {
        std::string n = std::to_string(unit);
        ag_write_file(tmp + "/_initial_contigs." + n + ".fa", r.initial_text);
        ag_write_file(tmp + "/_pre_extended_contigs." + n + ".fa", r.pre_text);
        ag_write_file(tmp + "/_extended_contigs." + n + ".fa", r.ext_text);
    }
}

// Node table as text, one line per node in (position, item) order — the format of the oracle's --dump-nodes, for node-level
// parity tests:  pos item cov A C G T N cid coff cid0 coff0 mid moff [s] succPos:succItem ...
inline void ag_format_node_dump(const AgNodeDump& d, const AgReads& reads, std::string& text) {
    text.clear();
    char buf[256];
    // global node index -> (pos, item)
    for (size_t v = 0; v < d.pos.size(); v++) {
        u32 moff = d.moff[v];
        int n = snprintf(buf, sizeof buf, "%u %u %u %u %u %u %u %u %u %u %u %u %u %u [", d.pos[v], d.item[v], d.cov[v], d.cnt[5 * v], d.cnt[5 * v + 1], d.cnt[5 * v + 2],
                         d.cnt[5 * v + 3], d.cnt[5 * v + 4], d.cid[v], d.coff[v], d.cid0[v], d.coff0[v], moff == AG_NONE ? AG_NONE : 0u, moff);
        text.append(buf, (size_t)n);
        u32 slen = d.soff_len[v] >> 16, soff = d.soff_len[v] & 0xFFFFu, read = d.sread[v] >> 1, rc = d.sread[v] & 1;
        u32 rlen = slen ? reads.len[read >> 1] : 0;
        for (u32 j = 0; j < slen; j++) text.push_back(reads.at(read, rc, rlen, soff + j));
        text.push_back(']');
        for (u32 e = d.edge_start[v]; e < d.edge_start[v + 1]; e++) {
            u32 t = d.edge_target[e];
            n = snprintf(buf, sizeof buf, " %u:%u", d.pos[t], d.item[t]);
            text.append(buf, (size_t)n);
        }
        text.push_back('\n');
    }
}

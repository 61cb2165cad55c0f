// Host-visible interface of the device pipeline (ag_device.cu).  Internal to the library; the public boundary is the C ABI in
// include/aligngraph_b200.h.
#pragma once
#include "ag_types.h"
#include <string>
#include <vector>
#include <functional>

struct AgUnitInput {
    const char* ref;            // unit bases followed by the contig-insertion tail (AG:981-1036), n_pos bytes
    u32 n_ref, n_pos;
    const ag_cthread* threads; u32 n_threads;   // contig threads: when given, the device derives cm_start / cm itself and they may be null
    const u32* cm_start;        // n_pos + 1
    const ag_cm* cm; u32 n_cm;
    const u32* chain_pos;       // n_cm, chain-major
    const char* chain_base;     // n_cm, chain-major
    const ag_aln* aln; u64 n_aln;
    const ag_seg* ext; u64 n_ext;
    // run-space contig threads (ag_thread_contigs_runs): when cdesc != null, chain_pos / chain_base are null and the device expands them
    const struct ag_cdesc* cdesc = nullptr; u32 n_desc = 0; const struct ag_crun* cruns = nullptr; u32 n_runs = 0;
    const char* contig_blob = nullptr; u64 blob_bytes = 0, blob_version = 0;
    bool aln_on_device = false;  // the alignment tuples were parsed on the device (AgDevice::ingest_sam): aln / ext are not uploaded
};

struct AgNodeDump {  // node table in (position, item) order, for tests
    std::vector<u32> pos, item, cov, cnt, cid, coff, cid0, coff0, moff, sread, soff_len;
    std::vector<u32> edge_start;   // CSR over nodes
    std::vector<u32> edge_target;  // global node index
};

struct AgTimings {  // milliseconds, CUDA events on the context's stream
    float h2d = 0, prep = 0, sort = 0, nodes = 0, finalize = 0, edges = 0, components = 0, chains = 0, walk = 0, materialize = 0, d2h = 0;
    u64 n_nodes = 0, n_edges_ovf = 0, n_walks = 0, n_keys = 0, n_tiles = 0, n_components = 0;
    u64 h2d_bytes = 0, d2h_bytes = 0;
    int walk_fallback = 0, regrows = 0;   // regrows: sweeps repeated with a larger node table / overflow pool
    float select = 0;                        // emission filter + materialisation inputs on the device (fused extension path)
    float stage = 0, build_kernel = 0;        // inside `nodes`: the staging gather (k_stage) and the node sweep kernel itself
    float ingest_reads = 0, ingest_sam = 0;   // ms: staging + kernels of the text ingestion (CUDA events)
    u64 sam_device = 0, sam_host = 0, reads_device = 0, reads_host = 0, reads_windowed = 0;   // files parsed on the device / by the host parser
};

class AgDevice {
public:
    explicit AgDevice(int device);
    ~AgDevice();
    // reads: 2-bit packed + non-ACGT bit plane, fixed stride per read; len per pair.  `on_device` = pointers are device pointers
    // (the NCCL broadcast target); otherwise they are copied H2D.
    void set_reads(const u32* bases, const u32* nmask, const uint16_t* len, u64 n_pairs, u32 stride2, u32 stridem, bool on_device);
    // the same from host memory, the non-ACGT plane given as its list of set bits (key = read * 65536 + offset): the plane is rebuilt here
    // `overlap`: the bulk of the copy (packed bases, exception list) goes on a second stream BEHIND whatever the main stream holds now
    // (the unit's own uploads), so that it runs under the unit's table / prep / bucket kernels; build() waits for it before the node sweep
    void set_reads_sparse(const u32* bases, const u64* exc_keys, u64 n_exc, const uint16_t* len, u64 n_pairs, u32 stride2, u32 stridem, bool overlap = false);
    void copy_reads_to_host(u32* bases, u32* nmask, uint16_t* len);
    // device-resident read set of this context as raw buffers (for the broadcast to the other GPUs of a run)
    struct ReadsView { u32* bases; u32* nmask; uint16_t* len; size_t bases_bytes, nmask_bytes, len_bytes; u64 n_pairs; u32 stride2, stridem; };
    ReadsView reads_view();
    // make own, uninitialised buffers of this geometry the context's read set (the caller fills them: broadcast target)
    ReadsView reserve_reads(u64 n_pairs, u32 stride2, u32 stridem);
    // ---- GPU-side text ingestion (ag_ingest.cuh).  Both return false when the file is not in the well-formed layout the kernels handle;
    // the caller then uses the sequential host parser (ag_parse_reads / ag_parse_sam), which carries the literal semantics and messages.
    // tmp/_reads.fa -> packed reads resident on the device; `host` receives the per-pair lengths, the geometry and the exception list
    // (original non-ACGT characters), NOT the packed words (copy_reads_to_host fetches them on demand)
    // win_lo / win_hi >= 0: only the pairs of that id window are made resident (their byte range of the file is found by bisection; ids are
    // the pair indices, AG:3455-3471) — ingest_sam then reports ids outside the window through sam_window_miss()
    bool ingest_reads(const std::string& path, struct AgReads& host, long long win_lo = -1, long long win_hi = -1);
    bool sam_window_miss() const;
    void set_reads_window(u64 lo, u64 hi);   // pair-id window of the resident read set (a broadcast target inherits the source's)
    // tmp/_reads_genome.N.bowtie -> the unit's surviving alignment tuples, resident on the device in file order
    bool ingest_sam(const std::string& path);
    // removeMisassembly's coverage pile-up (AG:3938-3978) of tmp/_reads_<id>_contigs.bowtie over the chunks of tmp/_<id>_contigs.fa; false: the
    // file is not in the layout the kernel handles (the caller uses ag_coverage_pileup_host)
    bool coverage_pileup(const std::string& sam_path, const std::vector<u32>& chunk_len, std::vector<int>& coverage);
    // matching bases of the candidate placements of ag_contain_search (refinement without BLAT), on the device
    void verify_placements(const struct AgSeqSet& db, const struct AgSeqSet& queries, const std::vector<struct AgPlacement>& cand, std::vector<u32>& match);
    u64 ingested_alignments() const;
    void note_host_sam() { t_.sam_host++; }
    void note_host_reads() { t_.reads_host++; }
    void fetch_alignments(std::vector<ag_aln>& aln, std::vector<ag_seg>& ext);   // device -> host copy of the ingested tuples (tests, ag_get_unit)
    void set_params(int k, int iv, int coverage);
    // keep coverage + base counters per node after the build (24 B per node; only the node dump of the tests needs them)
    void set_keep_counts(bool on) { keep_counts_ = on; }
    void set_option(const std::string& name, long value);
    // upload one unit's inputs (H2D, timed)
    void load_unit(const AgUnitInput& in);
    // the hot path: prep -> bucket -> nodes (+ common-case edges) -> successor lists -> generic edges on flagged tiles  (all device).
    // build() only QUEUES the kernels; errors and capacity overflows surface at the next synchronisation point — build_sync() (blocks,
    // checks, repeats the build with larger capacities if needed) or extend(), which queues the walk behind the build and synchronises once
    void build();
    void build_sync();
    // coverage filter + walk simulation; fills `walks` (unsorted on return from the device, sorted here by start node)
    void extend(std::vector<ag_walk>& walks);
    // extendContigs1 (AG:1954-2204) up to the emitted contigs as ONE queued step behind the build: walk, emission filter (a prefix maximum over
    // the walk records), materialisation and the copies to page-locked host memory; a single synchronisation.  `emitted` = the walk records that
    // pass the emission filter, in scan order; contig i = bases[offs[i], offs[i + 1]) (valid until the next call); the occupancy bitmap has been
    // queued as well (occupancy_wait)
    // on_records (optional) is called, possibly more than once (each call supersedes the earlier ones), as soon as `emitted` and `offs` are in host
    // memory and `bases` points to where the bases WILL be: the caller may build records and headers while the materialisation runs.  Returns
    // true when the last on_records call saw the final result (false: the caller starts from the returned arrays).
    bool extend_emitted(std::vector<ag_walk>& emitted, char*& bases, std::vector<u64>& offs, u64& n_walks, const std::function<void()>& on_records = nullptr);
    bool fused_extend() const { return !fused_off_; }
    // materialise the selected walks' base strings (loop bases + tail); contig i occupies bases[offs[i], offs[i + 1]).  `bases` points into
    // a page-locked buffer owned by the device object (valid until the next call); the post passes patch and read it in place
    void materialize(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char*& bases, std::vector<u64>& offs) { materialize_begin(walks, sel, bases, offs); materialize_wait(); }
    // the same in two halves: _begin queues the kernels and the copy and returns at once (`bases` / `offs` are final, the bytes are not
    // there yet), _wait blocks until they are
    void materialize_begin(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char*& bases, std::vector<u64>& offs);
    void materialize_wait();
    // occupancy bitmap (any node or contiMer at a position) for the scaffold gap test (AG:2428)
    void occupancy(std::vector<unsigned char>& bits) { occupancy_begin(); occupancy_wait(bits); }
    void occupancy_begin();
    void occupancy_wait(std::vector<unsigned char>& bits);
    void dump_nodes(AgNodeDump& d);
    void sync();
    void pin(const void* p, size_t bytes);
    void unpin_all();
    void timer_start();
    float timer_stop();
    const AgTimings& timings() const { return t_; }
    void reset_timings() { t_ = AgTimings(); launches_ = 0; }
    int device() const { return dev_; }
    void* stream() const { return stream_; }
    u64 kernel_launches() const { return launches_; }

private:
    struct Impl;
    Impl* m_;
    int dev_, k_ = 5, iv_ = 50, cov_ = 20;
    void* stream_ = nullptr;
    AgTimings t_;
    u64 launches_ = 0;
    void *ev0_ = nullptr, *ev1_ = nullptr;
    void *st2_ = nullptr, *ev_main_ = nullptr, *ev_reads_ = nullptr;   // copy stream of the overlapped reads upload
    bool reads_pending_ = false, mat_pending_ = false, occ_pending_ = false;
    void *ev_mat0_ = nullptr, *ev_mat1_ = nullptr; size_t mat_bytes_ = 0;
    std::vector<void*> pinned_;
    bool tma_off_ = false, attr_tma_done_ = false, fused_off_ = false;
    bool chains_valid_ = false, attr_done_ = false, keep_counts_ = false, section_timing_ = false;
    void enqueue_build();
    void enqueue_walk(bool records_to_host = true);
    void enqueue_select();
    bool finish();
    void walk_sequential();
};

// One broadcast of the packed read set from devs[0] to devs[1..n-1]: ncclBroadcast over NVLink when libnccl can be loaded and
// the devices are distinct, else cudaMemcpyPeerAsync / device-to-device copies.  Returns "nccl" or "peer-copy".
const char* ag_device_broadcast_reads(AgDevice** devs, int n, double* seconds, size_t* bytes);

// thrown on any CUDA failure or capacity error; the C ABI turns it into an error code + ag_last_error()
struct AgError { std::string msg; };

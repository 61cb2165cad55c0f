// Host side of the B200 AlignGraph hot path: the text <-> packed-array boundary of the reference's tmp/ file contract
//  and the order-dependent FASTA post-passes.  Everything here is cheap, sequential text / interval logic; the
// graph build and the walk run on the device (ag_device.cu).
#pragma once
#include "ag_types.h"
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct AgHostError { std::string msg; };

// Zero-initialised growable array backed by calloc: large blocks come from fresh (already zero) pages that are first touched by whoever
// writes them — the parallel packers — instead of being memset by one thread as std::vector would do.
template <class T> struct AgZVec {
    T* p = nullptr; size_t n = 0;
    AgZVec() {}
    AgZVec(const AgZVec& o) { assign(o.p, o.p + o.n); }
    AgZVec& operator=(const AgZVec& o) { if (this != &o) assign(o.p, o.p + o.n); return *this; }
    ~AgZVec() { free(p); }
    void assign_zero(size_t count) { free(p); p = count ? (T*)calloc(count, sizeof(T)) : nullptr; n = count; if (count && !p) throw AgHostError{"out of host memory"}; }
    void assign(const T* b, const T* e) { assign_zero((size_t)(e - b)); if (n) memcpy(p, b, n * sizeof(T)); }
    void resize(size_t count) { assign_zero(count); }
    T* data() { return p; } const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; } const T& operator[](size_t i) const { return p[i]; }
};  // message the CLI prints on stdout before exit(-1), as the reference does

// Text output buffer whose growth does not zero-fill (the formatters write every byte, several threads at once) and whose capacity
// survives from unit to unit.
struct AgText {
    char* p = nullptr; size_t n = 0, cap = 0;
    AgText() {}
    AgText(const AgText& o) { assign(o.p, o.n); }
    AgText& operator=(const AgText& o) { if (this != &o) assign(o.p, o.n); return *this; }
    AgText(AgText&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    AgText& operator=(AgText&& o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
    ~AgText() { free(p); }
    void set_size(size_t m) {   // contents undefined afterwards
        if (m > cap) { free(p); cap = m + m / 8 + 64; p = (char*)malloc(cap); if (!p) { cap = 0; throw AgHostError{"out of host memory"}; } }
        n = m;
    }
    void assign(const char* s, size_t m) { set_size(m); if (m) memcpy(p, s, m); }
    void clear() { n = 0; }
    const char* data() const { return p; } char* data() { return p; }
    size_t size() const { return n; } bool empty() const { return n == 0; }
    bool operator==(const std::string& s) const { return s.size() == n && (n == 0 || memcmp(s.data(), p, n) == 0); }
};
// run fn(chunk) for chunk = 0 .. n_chunks-1 on the host thread team (persistent workers + the caller; AG_THREADS / affinity-aware)
void ag_parallel_chunks(int n_chunks, const std::function<void(int)>& fn);
int ag_team_size();
// thread budget of the text parsers called from THIS thread (0 = default: AG_THREADS or all cores)
void ag_set_thread_budget(int n);

// ---- reads (tmp/_reads.fa, AG:361-404) -----------------------------------------------------------------------------------
struct AgReads {
    AgZVec<u32> bases, nmask;        // 2 bits / base, 16 per word; 1 bit / base, 32 per word; fixed stride per read
    std::vector<uint16_t> len;       // per pair
    std::vector<std::pair<u64, char>> exc;  // (read * 65536 + offset, original character) for every non-ACGT character, sorted
    bool exc_complete = false;       // exc lists EVERY set bit of nmask (true after ag_parse_reads / ag_pack_reads): the plane can be rebuilt from it
    u64 n_pairs = 0;
    u64 win_lo = 0, win_hi = ~0ull;   // pair-id window for which lengths / packed words are held (everything unless the set was ingested for a window)
    bool windowed() const { return win_hi != ~0ull && (win_lo != 0 || win_hi + 1 < n_pairs); }
    u32 stride2 = 0, stridem = 0;
    // character of the oriented read (AG:854-865 leaves non-ACGT unchanged).  When the packed words live only on the device (GPU-side
    // ingestion) the exception list alone is consulted: the original character of a masked base, or 0 = "an ordinary base"
    char at(u32 read, u32 rc, u32 rlen, u32 off) const;
    bool device_only() const { return n_pairs != 0 && bases.size() == 0; }
};
void ag_parse_reads(const std::string& path, AgReads& out);
// pack in-memory reads (one string per read, mates interleaved) — used by the synthetic bench and tests
void ag_pack_reads(const std::vector<std::string>& seqs, AgReads& out);

// ---- contig chunks (tmp/_contigs.fa, AG:322-359): parsed once per file version, shared by all units and contexts --------------
struct AgChunkStore {
    std::vector<int> id;        // original contig of every chunk (">chunk.contig")
    std::vector<u64> off;       // n + 1 offsets into blob
    std::string blob;           // chunk bases, concatenated
    u64 version = 0;            // changes whenever the file is re-parsed (device-side cache key)
    size_t n() const { return id.size(); }
    size_t size(size_t i) const { return (size_t)(off[i + 1] - off[i]); }
};
std::shared_ptr<const AgChunkStore> ag_chunk_store(const std::string& contigs_fa);

// One contig thread in run space (what updateGenomeWithContig, AG:884-1177, produces for one position set): the chunk's bases F .. L map
// to unit positions through `nruns` runs of aligned bases; bases in the gaps between runs are insertions and live in the tail appended
// behind the unit (AG:981-1036).  The thread's contiMers are chain indices [first, first + n): n - 1 bases F .. L-1 and the terminal
// contiMer at the position of base L (AG:1121-1148).  The device expands these descriptors into the chain-major position / base arrays.
struct ag_crun { u32 src, dst, len, gap_tail; };   // gap_tail: tail index (relative to n_ref) of the first inserted base that follows this run
struct ag_cdesc { u32 first, n, F, size, fr, run0, nruns, pad; u64 base_off; };   // base_off: offset of the chunk's bases in the store blob

// ---- one unit -----------------------------------------------------------------------------------------------------------------
struct AgUnit {
    std::string ref;                 // unit bases + contig-insertion tail
    u32 n_ref = 0;
    std::vector<ag_cthread> threads; // contig threads; when non-empty (or cm empty) the contiMer table below is derived from them
    std::vector<u32> cm_start;       // CSR over positions (ref.size() + 1) — only filled by ag_expand_contimers / ag_set_contimers
    std::vector<ag_cm> cm;
    std::vector<u32> chain_pos;      // chain-major
    std::string chain_base;
    std::vector<ag_aln> aln;
    std::vector<ag_seg> ext;
    // run-space form of the contig threads (ag_thread_contigs_runs): when cdesc is non-empty chain_pos / chain_base stay EMPTY on the host
    // and the device expands them (k_chain_expand); ag_expand_chains fills them on the host for inspection / the emulation
    std::vector<ag_cdesc> cdesc; std::vector<ag_crun> cruns; std::shared_ptr<const AgChunkStore> chunks; u32 n_cm_runs = 0;
    bool aln_on_device = false;      // the SAM was parsed on the device (AgDevice::ingest_sam): aln / ext stay empty here unless fetched for inspection
};
void ag_load_genome(const std::string& path, AgUnit& u);                                         // AG:287-320
// contig chunks + PSL -> contiMer table + the text of tmp/_initial_contigs.N.fa                       // AG:1219-1231
void ag_thread_contigs(const std::string& contigs_fa, const std::string& psl, std::string& initial_text, AgUnit& u);
// The same in run space: interval arithmetic on the PSL blocks instead of per-base position arrays (O(blocks) + one pass over a per-position
// counter).  Fills u.threads, the tail of u.ref, u.cdesc / u.cruns / u.chunks and initial_text; returns false (u untouched apart from ref's
// tail being reset) for inputs it does not model (a record whose own blocks overlap) — the caller then uses ag_thread_contigs.
bool ag_thread_contigs_runs(const std::string& contigs_fa, const std::string& psl, std::string& initial_text, AgUnit& u);
// chain_pos / chain_base of a run-space unit on the host (what the device derives)
void ag_expand_chains(AgUnit& u);
// position-ordered contiMer table (cm_start, cm) from the contig threads — what the device derives itself; host callers that want to
// look at the table (tests, ag_get_unit) call this
void ag_expand_contimers(AgUnit& u);
void ag_expand_contimers(const ag_cthread* threads, size_t n_threads, const u32* chain_pos, size_t n_cm, size_t n_pos, std::vector<u32>& cm_start, std::vector<ag_cm>& cm);
// SAM -> surviving alignments in processing order                                                 // AG:1233-1277, 1644-1656, 1872-1895
void ag_parse_sam(const std::string& path, const AgReads& reads, AgUnit& u);
bool ag_selfcheck_sam_line(const char* s, size_t n);   // tests: ag_samcore.h record parser == general record parser on this line

// ---- post passes --------------------------------------------------------------------------------------------------------------------
struct AgPiece { const char* p; size_t n; };
struct AgContig {
    int extended; u32 sid, soff, eid, eoff, sid0, soff0, eid0, eoff0;
    const char* p = nullptr; size_t n = 0;  // bases: a view into the materialised buffer ...
    std::vector<AgPiece> more;              // ... followed by views of the contigs it has been joined with (AG:2368-2370); nothing is copied
    size_t size() const { size_t t = n; for (const AgPiece& q : more) t += q.n; return t; }
    // append the pieces of this contig from byte offset `from` on to `out`
    void pieces_from(size_t from, std::vector<AgPiece>& out) const {
        if (from < n) out.push_back(AgPiece{p + from, n - from});
        size_t skip = from > n ? from - n : 0;
        for (const AgPiece& q : more) { if (skip >= q.n) { skip -= q.n; continue; } out.push_back(AgPiece{q.p + skip, q.n - skip}); skip = 0; }
    }
};
// emission filter of extdContigs1 (AG:2176-2189): indices of the walks that are written to _pre_extended_contigs
void ag_select_emitted(const std::vector<ag_walk>& walks, std::vector<u32>& sel);
// emitted contigs as views into `bases` (loop bases + s[1..] tail per walk, written by the device; characters outside ACGT are
// restored here from the reads' exception list) + the text of tmp/_pre_extended_contigs.N.fa
void ag_make_contigs(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                     const AgReads& reads, std::vector<AgContig>& contigs, AgText& pre_text);
// the same in two halves, so that the first (contig records, header lines, text offsets — nothing that reads the bases) can run while the
// device is still materialising and copying them
struct AgMakeState { std::string hdr; std::vector<size_t> hoff, toff; };
void ag_make_contigs_begin(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                           std::vector<AgContig>& contigs, AgMakeState& st);
void ag_make_contigs_finish(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                            const AgReads& reads, const std::vector<AgContig>& contigs, const AgMakeState& st, AgText& pre_text);
void ag_dedup_join(std::vector<AgContig>& contigs);                                               // AG:2296-2380
// AG:2396-2464; occ = bitmap "position holds a node or a contiMer"
void ag_scaffold(std::vector<AgContig>& contigs, const std::string& ref, const std::vector<unsigned char>& occ, AgText& text);

// ---- input normalisation re-run by --resume (AG:3228-3345, AG:3347-3418) ----------------------------------------------------
void ag_formalize_contigs(const std::string& in_path, const std::string& tmp, std::vector<std::string>& contig_ids);
int ag_formalize_genome(const std::string& in_path, const std::string& tmp, int part, std::vector<std::string>& genome_ids);
void ag_write_file(const std::string& path, const std::string& text);
void ag_write_file(const std::string& path, const AgText& text);
// glibc tuning for the staging buffers (see ag_host.cpp); no-op when AG_NO_MALLOPT is set
void ag_tune_malloc();

// ---- CLI phases outside the hot path (drop-in compatibility) ----------------------------------------------------------------
int ag_max_read_length(const std::string& path);                                                        // AG:3197
long ag_formalize_reads(const std::string& in1, const std::string& in2, const std::string& tmp);        // AG:3420
void ag_distribute_alignments(const std::string& tmp, int units);                                       // AG:3545
double ag_check_ratio(const std::string& tmp, int units);                                               // AG:3751
// Built-in replacement for the aligner call of refinement() (AG:2957-2983: `pblat <extended contigs> <truncated initial contigs> -noHead out.psl`)
// when no BLAT is installed: ungapped seed-and-verify containment search.  Every query (<= 20 kbp by construction, AG:2891-2953) is placed on
// every database sequence, either strand, by exact 24-mer seeds and a full-length comparison (>= 90 % identity: refinement only keeps
// alignments covering >= 80 % of the query, AG:3059); a query without any full-length placement gets its local ungapped alignments
// (seed + X-drop).  Output: single-block PSL lines (-noHead), one per placement.  `verify` counts the matching bases of every candidate
// placement — the device kernel in the product, ag_verify_placements_host in CPU tests.
struct AgPlacement { u32 q, strand, t; long start; };
struct AgSeqSet { std::vector<std::string> names; std::vector<u64> off; std::string blob; size_t n() const { return names.size(); } size_t len(size_t i) const { return (size_t)(off[i + 1] - off[i]); } };
typedef void (*AgVerifyFn)(const AgSeqSet& db, const AgSeqSet& queries, const std::vector<AgPlacement>& cand, std::vector<u32>& match, void* user);
void ag_verify_placements_host(const AgSeqSet& db, const AgSeqSet& queries, const std::vector<AgPlacement>& cand, std::vector<u32>& match, void* user);
void ag_contain_search(const std::string& db_fa, const std::string& query_fa, const std::string& out_psl, AgVerifyFn verify, void* user);

// removeMisassembly (AG:3821-4297) for one output file (`id` = "extended" | "remaining"): formalize it into tmp/_<id>_contigs.fa, let `align`
// run the aligners (bowtie2 reads -> contigs, BLAT contigs -> genome: command lines in the CLI), pile the read alignments up into a per-base
// coverage (`pileup`: the device kernel in the product, ag_coverage_pileup_host in CPU tests), keep / break / drop contig regions, write
// corrected_<file>.
typedef bool (*AgAlignFn)(const std::string& id, void* user);
typedef void (*AgPileupFn)(const std::string& sam_path, const std::vector<u32>& chunk_len, std::vector<int>& coverage, void* user);
void ag_remove_misassembly(const std::string& file, const std::string& id, int coverage, const std::string& tmp, AgAlignFn align, AgPileupFn pileup, void* user);
// per-base read coverage of the chunks of tmp/_<id>_contigs.fa from tmp/_reads_<id>_contigs.bowtie (AG:3923-3970), sequential host version
void ag_coverage_pileup_host(const std::string& sam_path, const std::vector<u32>& chunk_len, std::vector<int>& coverage, void* user);
void ag_formalize_contigs_to(const std::string& in_path, const std::string& out_path, const std::string& chaff_path /* empty: none */, std::vector<std::string>& contig_ids);
void ag_refinement(const std::string& tmp, int units, const std::vector<std::string>& genome_ids, const std::vector<std::string>& contig_ids,
                   int unique_extension, const std::string& ext_path, const std::string& rmn_path, bool (*blat)(int unit, void* user), void* user,
                   bool write_test_files);                                                              // AG:2864

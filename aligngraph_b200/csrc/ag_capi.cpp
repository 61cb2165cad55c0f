// C ABI of the B200 AlignGraph hot path (include/aligngraph_b200.h).  Thin: owns the context, stages host arrays, turns
// exceptions into error codes.  All graph work happens in AgDevice (ag_device.cu).
#include "../../include/aligngraph_b200.h"
#include "ag_pipeline.h"
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <map>
#include <atomic>
#include <memory>

static_assert(sizeof(ag_aln_c) == sizeof(ag_aln) && sizeof(ag_seg_c) == sizeof(ag_seg) && sizeof(ag_cm_c) == sizeof(ag_cm), "ABI structs mirror the internal layouts");

struct ag_ctx {
    AgDevice* dev = nullptr;
    ag_params params{};
    std::shared_ptr<AgReads> rp = std::make_shared<AgReads>();   // host copy of the read set; shared by the contexts of one run (ag_broadcast_reads)
    bool have_reads = false, reads_dirty = false;   // reads_dirty: a re-upload was requested; it is issued by the next ag_build, behind the unit's uploads
    AgUnit unit;
    int unit_id = -1;
    bool uploaded = false;  // the staged unit arrays are resident on the device
    AgUnitResult res;
    std::string err, dump;
    std::vector<u64> exc_keys;
    double s_parse = 0, s_device = 0, s_post = 0;
    u64 n_aln = 0, n_walks = 0, n_emitted = 0;
    bool host_parse = getenv("AG_HOST_PARSE") != nullptr;
    bool reads_window = getenv("AG_NO_READS_WINDOW") == nullptr;   // job-level loads may restrict the resident read set to the ids the job's SAM files reference
    std::string reads_path;                                         // where the resident read set came from (to load the rest when a window turns out too small)
    bool fused = false;
};

static std::string g_create_error;
extern "C" { static void load_reads_impl(ag_ctx* ctx, const char* path, long long win_lo, long long win_hi); }

static bool device_ingest_enabled(const ag_ctx* ctx) { return !ctx->host_parse; }   // AG_HOST_PARSE=1 / option "host_parse": text is parsed by the host parsers only

// the packed words of a device-ingested read set, fetched from the device when a host-side consumer asks for them
static void ensure_host_reads(ag_ctx* ctx) {
    AgReads& r = *ctx->rp;
    if (!r.device_only()) return;
    r.bases.resize(2 * r.n_pairs * r.stride2); r.nmask.resize(2 * r.n_pairs * r.stridem);
    std::vector<uint16_t> len(r.n_pairs);
    ctx->dev->copy_reads_to_host(r.bases.data(), r.nmask.data(), len.data());
}
// SAM of unit `unit_id`: parsed on the device (the tuples stay there); files outside the well-formed layout go to the host parser
static void load_unit_sam(ag_ctx* ctx, const std::string& tmp, int unit_id) {
    AgUnit& u = ctx->unit;
    const std::string path = tmp + "/_reads_genome." + std::to_string(unit_id) + ".bowtie";
    u.aln.clear(); u.ext.clear(); u.aln_on_device = false;
    if (device_ingest_enabled(ctx) && ctx->dev->ingest_sam(path)) { u.aln_on_device = true; return; }
    if (ctx->rp->windowed()) {   // only a window of the read set is resident and this file needs more (or the host parser
        if (ctx->reads_path.empty()) throw AgHostError{"reads not set"};    // does, which looks up any read's length): load the whole set, then try again
        load_reads_impl(ctx, ctx->reads_path.c_str(), -1, -1);
        if (device_ingest_enabled(ctx) && ctx->dev->ingest_sam(path)) { u.aln_on_device = true; return; }
    }
    ag_parse_sam(path, *ctx->rp, u);
    ctx->dev->note_host_sam();
}

template <class F> static int guard(ag_ctx* ctx, F f) {
    try { f(); return 0; }
    catch (const AgHostError& e) { if (ctx) ctx->err = e.msg; return 1; }
    catch (const AgError& e) { if (ctx) ctx->err = e.msg; return 2; }
    catch (const std::bad_alloc&) { if (ctx) ctx->err = "out of host memory"; return 3; }
    catch (...) { if (ctx) ctx->err = "unknown error"; return 4; }
}

extern "C" {

int ag_create(const ag_params* params, ag_ctx** out) {
    if (!params || !out) { g_create_error = "null argument"; return 1; }
    if (getenv("AG_MALLOPT")) ag_tune_malloc();   // process-wide allocator policy: opt-in for library users (the CLI applies it in its own main)
    ag_ctx* c = new (std::nothrow) ag_ctx;
    if (!c) { g_create_error = "out of host memory"; return 3; }
    c->params = *params;
    try { c->dev = new AgDevice(params->device); c->dev->set_params(params->k, params->insert_variation, params->coverage); }
    catch (const AgError& e) { g_create_error = e.msg; delete c; return 2; }
    *out = c;
    return 0;
}
void ag_destroy(ag_ctx* ctx) { if (!ctx) return; if (ctx->dev) ctx->dev->unpin_all(); delete ctx->dev; delete ctx; }
const char* ag_last_error(const ag_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
const char* ag_create_error(void) { return g_create_error.c_str(); }

// host -> device copy of the context's packed reads.  The non-ACGT bit plane is all zeros except where the parser recorded an exception,
// so when that list is complete only the list travels (8 bytes per non-ACGT character) and the device rebuilds the plane.
static void upload_reads(ag_ctx* ctx, bool overlap = false) {
    ensure_host_reads(ctx);
    const AgReads& r = (*ctx->rp);
    if (r.exc_complete && r.exc.size() * 8 < r.nmask.size() * 4) {
        ctx->exc_keys.resize(r.exc.size());
        for (size_t i = 0; i < r.exc.size(); i++) ctx->exc_keys[i] = r.exc[i].first;
        ctx->dev->set_reads_sparse(r.bases.data(), ctx->exc_keys.data(), ctx->exc_keys.size(), r.len.data(), r.n_pairs, r.stride2, r.stridem, overlap);
    } else ctx->dev->set_reads(r.bases.data(), r.nmask.data(), r.len.data(), r.n_pairs, r.stride2, r.stridem, false);
    ctx->reads_dirty = false;
}

int ag_set_reads(ag_ctx* ctx, const uint32_t* bases2, const uint32_t* nmask, const uint16_t* pair_len, uint64_t n_pairs, uint32_t stride2, uint32_t stridem) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all();
        ctx->rp = std::make_shared<AgReads>();   // detach from any run-wide shared copy
        AgReads& r = (*ctx->rp);
        r.n_pairs = n_pairs; r.stride2 = stride2; r.stridem = stridem;
        r.bases.assign((const u32*)bases2, (const u32*)bases2 + 2 * n_pairs * stride2); r.nmask.assign((const u32*)nmask, (const u32*)nmask + 2 * n_pairs * stridem); r.len.assign(pair_len, pair_len + n_pairs);
        r.exc.clear(); r.exc_complete = false;
        ctx->dev->set_reads(bases2, nmask, pair_len, n_pairs, stride2, stridem, false);
        ctx->have_reads = true;
    });
}
int ag_set_reads_device(ag_ctx* ctx, const uint32_t* d_bases2, const uint32_t* d_nmask, const uint16_t* d_pair_len, uint64_t n_pairs, uint32_t stride2, uint32_t stridem) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all();
        ctx->rp = std::make_shared<AgReads>();
        AgReads& r = (*ctx->rp);
        r.n_pairs = n_pairs; r.stride2 = stride2; r.stridem = stridem;
        r.bases.resize(2 * n_pairs * stride2); r.nmask.resize(2 * n_pairs * stridem); r.len.resize(n_pairs); r.exc.clear(); r.exc_complete = false;
        ctx->dev->set_reads(d_bases2, d_nmask, d_pair_len, n_pairs, stride2, stridem, true);
        ctx->dev->copy_reads_to_host(r.bases.data(), r.nmask.data(), r.len.data());
        ctx->have_reads = true;
    });
}
int ag_set_read_exceptions(ag_ctx* ctx, const uint64_t* keys, const char* chars, uint64_t n) {
    return guard(ctx, [&] { (*ctx->rp).exc.clear(); (*ctx->rp).exc_complete = false; for (uint64_t i = 0; i < n; i++) (*ctx->rp).exc.push_back({keys[i], chars[i]}); });
}
int ag_load_reads_fasta(ag_ctx* ctx, const char* path) {
    return guard(ctx, [&] { load_reads_impl(ctx, path, -1, -1); });
}
static void load_reads_impl(ag_ctx* ctx, const char* path, long long win_lo, long long win_hi) {
    {
        ctx->dev->unpin_all();
        ctx->reads_path = path;
        auto t0 = std::chrono::steady_clock::now();
        ctx->rp = std::make_shared<AgReads>();
        if (device_ingest_enabled(ctx) && ctx->dev->ingest_reads(path, *ctx->rp, win_lo, win_hi)) {   // raw text -> device, packed there (ag_ingest.cuh)
            ctx->s_parse += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            ctx->have_reads = true; ctx->reads_dirty = false;
            return;
        }
        ctx->rp = std::make_shared<AgReads>();
        ag_parse_reads(path, (*ctx->rp));
        ctx->s_parse += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        upload_reads(ctx);
        ctx->dev->note_host_reads();
        ctx->have_reads = true;
    }
}
int ag_get_reads(ag_ctx* ctx, const uint32_t** bases2, const uint32_t** nmask, const uint16_t** pair_len, uint64_t* n_pairs, uint32_t* stride2, uint32_t* stridem) {
    return guard(ctx, [&] {
        ensure_host_reads(ctx);
        const AgReads& r = (*ctx->rp);
        *bases2 = r.bases.data(); *nmask = r.nmask.data(); *pair_len = r.len.data(); *n_pairs = r.n_pairs; *stride2 = r.stride2; *stridem = r.stridem;
    });
}

// One read set for all the contexts of a run: ctxs[0] has parsed / received it; the others share its host copy (exception list included,
// so that contig tails print the original non-ACGT characters whichever GPU ran the unit, AG:2167) and receive the packed device buffers
// through ONE broadcast (ncclBroadcast over NVLink).
int ag_broadcast_reads(ag_ctx** ctxs, int n_ctx, ag_bcast_info* info) {
    if (!ctxs || n_ctx <= 0 || !ctxs[0]) return 1;
    return guard(ctxs[0], [&] {
        if (!ctxs[0]->have_reads) throw AgHostError{"reads not set"};
        std::vector<AgDevice*> devs;
        for (int i = 0; i < n_ctx; i++) { if (!ctxs[i]) throw AgHostError{"null context"}; devs.push_back(ctxs[i]->dev); }
        for (int i = 1; i < n_ctx; i++) { ctxs[i]->dev->unpin_all(); ctxs[i]->rp = ctxs[0]->rp; ctxs[i]->have_reads = true; ctxs[i]->reads_dirty = false; ctxs[i]->reads_path = ctxs[0]->reads_path; }
        double s = 0; size_t b = 0;
        const char* how = ag_device_broadcast_reads(devs.data(), n_ctx, &s, &b);
        for (int i = 1; i < n_ctx; i++) ctxs[i]->dev->set_reads_window(ctxs[0]->rp->win_lo, ctxs[0]->rp->win_hi);
        if (info) { info->seconds = s; info->bytes = b; info->nccl = strcmp(how, "nccl") == 0; }
    });
}

int ag_begin_unit(ag_ctx* ctx, int unit_id, const char* ref_bases, uint32_t n_ref) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all(); ctx->unit = AgUnit(); ctx->res.reset(); ctx->unit_id = unit_id; ctx->uploaded = false;
        ctx->unit.ref.assign(ref_bases, n_ref); ctx->unit.n_ref = n_ref;
        ctx->unit.cm_start.assign((size_t)n_ref + 1, 0);
    });
}
int ag_set_contimers(ag_ctx* ctx, const uint32_t* cm_start, const ag_cm_c* cm, uint32_t n_cm, const uint32_t* chain_pos, const char* chain_base,
                     const char* tail_bases, uint32_t n_tail) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all();
        AgUnit& u = ctx->unit; ctx->uploaded = false;
        u.ref.resize(u.n_ref);
        if (n_tail) u.ref.append(tail_bases, n_tail);
        u.threads.clear();
        u.cm_start.assign(cm_start, cm_start + u.ref.size() + 1);
        u.cm.assign((const ag_cm*)cm, (const ag_cm*)cm + n_cm);
        u.chain_pos.assign(chain_pos, chain_pos + n_cm); u.chain_base.assign(chain_base, n_cm);
    });
}
int ag_set_contig_threads(ag_ctx* ctx, const ag_cthread_c* threads, uint32_t n_threads, const uint32_t* chain_pos, const char* chain_base, uint32_t n_cm,
                          const char* tail_bases, uint32_t n_tail) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all();
        AgUnit& u = ctx->unit; ctx->uploaded = false;
        u.ref.resize(u.n_ref);
        if (n_tail) u.ref.append(tail_bases, n_tail);
        u.threads.assign((const ag_cthread*)threads, (const ag_cthread*)threads + n_threads);
        u.cm_start.clear(); u.cm.clear();
        u.chain_pos.assign(chain_pos, chain_pos + n_cm); u.chain_base.assign(chain_base, n_cm);
        for (uint32_t i = 0; i < n_threads; i++)
            if (u.threads[i].first > u.threads[i].term || u.threads[i].term >= n_cm || (i && u.threads[i].first != u.threads[i - 1].term + 1) || (!i && u.threads[i].first != 0))
                throw AgHostError{"CONTIG ALIGNMENT ERROR"};
        if (n_threads ? u.threads.back().term + 1 != n_cm : n_cm != 0) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
    });
}
int ag_add_alignments(ag_ctx* ctx, const ag_aln_c* aln, uint64_t n, const ag_seg_c* ext, uint64_t n_ext) {
    return guard(ctx, [&] {
        ctx->dev->unpin_all();
        AgUnit& u = ctx->unit; ctx->uploaded = false;
        u32 base = (u32)u.ext.size();
        size_t a0 = u.aln.size();
        u.aln.insert(u.aln.end(), (const ag_aln*)aln, (const ag_aln*)aln + n);
        if (base) for (size_t i = a0; i < u.aln.size(); i++) u.aln[i].ext_idx += base;
        u.ext.insert(u.ext.end(), (const ag_seg*)ext, (const ag_seg*)ext + n_ext);
    });
}

int ag_build(ag_ctx* ctx) {
    return guard(ctx, [&] {
        if (!ctx->have_reads) throw AgHostError{"reads not set"};
        auto t0 = std::chrono::steady_clock::now();
        if (!ctx->uploaded) { ctx->dev->load_unit(ag_unit_input(ctx->unit)); ctx->uploaded = true; }
        if (ctx->reads_dirty) upload_reads(ctx, true);   // overlaps the unit's table / prep / bucket kernels
        ctx->dev->build();
        if (!ctx->fused) ctx->dev->build_sync();   // the stand-alone call reports the build's own errors; the fused step synchronises once, in ag_extend
        ctx->s_device += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ctx->n_aln += ctx->unit.aln_on_device ? ctx->dev->ingested_alignments() : ctx->unit.aln.size();
    });
}
// graph build + extension as ONE step: the build is only queued, the walk is queued behind it, and the host waits once (ag_extend)
int ag_process(ag_ctx* ctx) {
    ctx->fused = true;
    int rc = ag_build(ctx);
    ctx->fused = false;
    if (!rc) rc = ag_extend(ctx);
    return rc;
}

// extendContigs + scaffoldContigs on the device engine's fused path: walk, emission filter and materialisation are one queued step with a
// single synchronisation (AgDevice::extend_emitted); the host continues with the emitted contigs only
static void extend_unit_fused(AgDevice& dev, const AgReads& reads, const std::string& ref, AgUnitResult& r, u64& n_walks, u64& n_emitted) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<ag_walk> em; char* bases = nullptr; std::vector<u64> offs;
    std::vector<u32> sel;
    std::vector<AgContig> contigs; AgMakeState ms;
    auto begin = [&] {   // contig records + headers: needs the emitted walk records and offsets, not the bases
        sel.resize(em.size());
        for (size_t i = 0; i < sel.size(); i++) sel[i] = (u32)i;
        ag_make_contigs_begin(em, sel, bases, offs, contigs, ms);
    };
    const bool begun = dev.extend_emitted(em, bases, offs, n_walks, begin);   // (begin runs while the device is still producing the bases)
    auto t1 = std::chrono::steady_clock::now();
    if (!begun) begin();
    ag_make_contigs_finish(em, sel, bases, offs, reads, contigs, ms, r.pre_text);
    ag_dedup_join(contigs);
    std::vector<unsigned char> occ;
    dev.occupancy_wait(occ);
    ag_scaffold(contigs, ref, occ, r.ext_text);
    auto t2 = std::chrono::steady_clock::now();
    r.t_device += std::chrono::duration<double>(t1 - t0).count();
    r.t_post += std::chrono::duration<double>(t2 - t1).count();
    n_emitted = em.size();
}

// extension half of ag_process_unit (kept separate so that ag_build can be timed / inspected on its own)
int ag_extend(ag_ctx* ctx) {
    return guard(ctx, [&] {
        AgUnitResult& r = ctx->res;
        const double d0 = r.t_device, p0 = r.t_post;
        u64 nw = 0, ne = 0;
        if (ctx->dev->fused_extend()) extend_unit_fused(*ctx->dev, (*ctx->rp), ctx->unit.ref, r, nw, ne);
        else ag_extend_unit(*ctx->dev, (*ctx->rp), ctx->unit.ref, r, nw, ne);
        ctx->s_device += r.t_device - d0; ctx->s_post += r.t_post - p0;
        ctx->n_walks += nw; ctx->n_emitted += ne;
    });
}

int ag_get_text(ag_ctx* ctx, int which, const char** text, uint64_t* len) {
    return guard(ctx, [&] {
        const char* tp = which == 0 ? ctx->res.initial_text.data() : which == 1 ? ctx->res.pre_text.data() : ctx->res.ext_text.data();
        const size_t tn = which == 0 ? ctx->res.initial_text.size() : which == 1 ? ctx->res.pre_text.size() : ctx->res.ext_text.size();
        *text = tp; *len = tn;
    });
}

int ag_prepare_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id) {
    return guard(ctx, [&] {
        if (!ctx->have_reads) throw AgHostError{"reads not set"};
        auto t0 = std::chrono::steady_clock::now();
        ctx->dev->unpin_all(); ctx->unit = AgUnit(); ctx->res.reset(); ctx->unit_id = unit_id; ctx->uploaded = false;
        ag_prepare_unit((*ctx->rp), tmp_dir, unit_id, ctx->unit, ctx->res.initial_text, false, ctx->host_parse);
        load_unit_sam(ctx, tmp_dir, unit_id);
        ctx->s_parse += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}
int ag_write_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id) {
    return guard(ctx, [&] {
        std::string n = std::to_string(unit_id), tmp = tmp_dir;
        std::string errs[3];
        ag_parallel_chunks(3, [&](int i) {   // the three files side by side
            try {
                if (i == 0) ag_write_file(tmp + "/_initial_contigs." + n + ".fa", ctx->res.initial_text);
                else if (i == 1) ag_write_file(tmp + "/_pre_extended_contigs." + n + ".fa", ctx->res.pre_text);
                else ag_write_file(tmp + "/_extended_contigs." + n + ".fa", ctx->res.ext_text);
            } catch (const AgHostError& e) { errs[i] = e.msg; }
        });
        for (const std::string& e : errs) if (!e.empty()) throw AgHostError{e};
    });
}
int ag_run_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id) {
    int rc = ag_prepare_unit_files(ctx, tmp_dir, unit_id);
    if (!rc) rc = ag_process(ctx);
    if (!rc) rc = ag_write_unit_files(ctx, tmp_dir, unit_id);
    return rc;
}

// Several units through the file-level path with the host half pipelined: `prefetch` threads parse units ahead (SAM / PSL / FASTA ->
// packed arrays) while one worker thread per context uploads, builds, extends and writes.  Units are handed out in order; outputs do
// not depend on the number of contexts or on timing (units are independent, AG:4779-4781).  `done` is called (serialised) after every
// unit with its return code; the first failing unit's code is returned.
// first and last read id of a SAM file (QNAME of the first record after the '@' lines and of the last line); false when there is no record
static bool sam_id_range(const std::string& path, long long& lo, long long& hi) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    char buf[1 << 16];
    auto id_of = [](const char* s, const char* e, long long& v) { v = 0; const char* p = s; bool any = false; while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); p++; any = true; } return any && p < e && *p == '\t'; };
    size_t got = fread(buf, 1, sizeof buf, f);
    const char* p = buf; const char* e = buf + got;
    while (p < e && *p == '@') { const char* l = (const char*)memchr(p, '\n', (size_t)(e - p)); if (!l) { p = e; break; } p = l + 1; }
    if (p < e && id_of(p, e, lo)) {
        if (fseek(f, 0, SEEK_END) == 0) {
            const long n = ftell(f); const long w = n < (long)sizeof buf ? n : (long)sizeof buf;
            if (fseek(f, n - w, SEEK_SET) == 0 && (got = fread(buf, 1, (size_t)w, f)) == (size_t)w && w > 1 && buf[w - 1] == '\n') {
                long i = w - 2;
                while (i >= 0 && buf[i] != '\n') i--;
                if ((i >= 0 || w == n) && id_of(buf + i + 1, buf + w, hi)) ok = true;
            }
        }
    }
    fclose(f);
    return ok;
}

// only the reads a set of units can reference need to be resident: the id window spanned by the first and last records of their SAM files (files
// are in read order, AG:3602 --reorder); a record outside it is noticed when its SAM is parsed and the whole set is loaded then
static void units_read_window(const std::string& tmp, const std::vector<int>& units, long long& wlo, long long& whi) {
    wlo = whi = -1;
    for (int u : units) { long long a, b; if (sam_id_range(tmp + "/_reads_genome." + std::to_string(u) + ".bowtie", a, b)) { if (a > b) std::swap(a, b); wlo = wlo < 0 ? a : std::min(wlo, a); whi = std::max(whi, b); } }
    if (wlo < 0) wlo = whi = 0;   // no record anywhere: nothing will be looked up
}
int ag_load_reads_for_units(ag_ctx* ctx, const char* reads_fa, const char* tmp_dir, const int* units, int n_units) {
    return guard(ctx, [&] {
        long long wlo = -1, whi = -1;
        if (ctx->reads_window && !ctx->host_parse && n_units > 0) units_read_window(tmp_dir, std::vector<int>(units, units + n_units), wlo, whi);
        load_reads_impl(ctx, reads_fa, wlo, whi);
    });
}

// The whole hot loop for a list of units: [reads] -> per unit (host: genome + contig threads; device: SAM, graph build, walk; host: post passes,
// files).  `prefetch` host threads prepare units ahead (in list order) while one worker thread per context runs them; with reads_fa != NULL
// the read set is (re)loaded as part of the job — raw text to ctxs[0]'s GPU, one broadcast to the others — while the preparers already work
// on the first units.  Outputs do not depend on the number of contexts or on timing (units are independent, AG:4779-4781).
static int run_units_impl(ag_ctx** ctxs, int n_ctx, const char* tmp_dir, const std::vector<int>& units, int prefetch, const char* reads_fa,
                          void (*done)(int unit, int rc, const char* error, void* user), void* user) {
    const int n_units = (int)units.size();
    if (!ctxs || n_ctx <= 0) return 1;
    for (int i = 0; i < n_ctx; i++) if (!ctxs[i]) return 1;
    if (!reads_fa) for (int i = 0; i < n_ctx; i++) if (!ctxs[i]->have_reads) { ctxs[i]->err = "reads not set"; return 1; }
    if (n_units == 0 && !reads_fa) return 0;
    struct Prepared { AgUnit unit; std::string initial_text, error; double s_parse = 0; };
    std::mutex mu; std::condition_variable cv;
    std::map<int, std::unique_ptr<Prepared>> ready;   // by position in `units`
    std::atomic<int> next_prepare(0), next_run(0);
    int consumed = 0;   // list positions [0, consumed) have been taken by workers (bounds the look-ahead)
    if (prefetch < 1) prefetch = 1;
    const std::string tmp = tmp_dir;
    const bool host_sam = !device_ingest_enabled(ctxs[0]);
    std::atomic<bool> stop(false);   // a unit failed: no further units are started (the reference exits at the failing chromosome)
    const int n_prep = std::min(prefetch, std::max(n_units, 1));
    const int cores = std::max(1u, std::thread::hardware_concurrency());
    auto preparer = [&]() {
        // The preparers run side by side.  Measured on the 16-core GPU host (profiles/r02g_scale_c4_1gpu.json): letting every parser start its
        // full set of threads keeps the cores busy through the others' serial phases; a strict share-out (cores / n_prep each) is available
        // through AG_PREP_THREADS.
        if (const char* e = getenv("AG_PREP_THREADS")) { int n = atoi(e); if (n > 0) ag_set_thread_budget(n); else if (n < 0) ag_set_thread_budget(std::max(1, (cores + n_prep - 1) / n_prep)); }
        for (;;) {
            int i = next_prepare.fetch_add(1);
            if (i >= n_units) return;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return i < consumed + prefetch + n_ctx || stop.load(); }); }
            if (stop.load()) return;
            auto p = std::make_unique<Prepared>();
            auto t0 = std::chrono::steady_clock::now();
            try { ag_prepare_unit(*ctxs[0]->rp, tmp, units[(size_t)i], p->unit, p->initial_text, host_sam, host_sam); }
            catch (const AgHostError& e) { p->error = e.msg; }
            catch (const std::bad_alloc&) { p->error = "out of host memory"; }
            p->s_parse = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            { std::lock_guard<std::mutex> lk(mu); ready[i] = std::move(p); }
            cv.notify_all();
        }
    };
    int first_rc = 0;
    // AG_JOB_TIMING: wall-clock laps of the job's serial path on stderr (diagnosis)
    const bool job_timing = getenv("AG_JOB_TIMING") != nullptr;
    auto T_job = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (job_timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  [job] %-32s %7.2f ms\n", what, std::chrono::duration<double>(t - T_job).count() * 1e3); T_job = t; } };
    auto worker = [&](ag_ctx* ctx) {
        for (;;) {
            int i = next_run.fetch_add(1);
            if (i >= n_units || stop.load()) return;
            const int u = units[(size_t)i];
            std::unique_ptr<Prepared> p;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return ready.count(i) != 0 || stop.load(); }); if (!ready.count(i)) return; p = std::move(ready[i]); ready.erase(i); if (i + 1 > consumed) consumed = i + 1; }
            cv.notify_all();
            int rc = 0;
            lap("wait for the prepared unit");
            if (!p->error.empty()) { ctx->err = p->error; rc = 1; }
            else {
                ctx->dev->unpin_all();
                ctx->unit = std::move(p->unit); ctx->res.reset(); ctx->res.initial_text = std::move(p->initial_text);
                ctx->unit_id = u; ctx->uploaded = false; ctx->s_parse += p->s_parse;
                lap("unit handed over");
                if (!host_sam) rc = guard(ctx, [&] { auto t0 = std::chrono::steady_clock::now(); load_unit_sam(ctx, tmp, u); ctx->s_parse += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); });
                lap("SAM text -> device tuples");
                if (!rc) rc = ag_process(ctx);
                lap("process");
                if (!rc) rc = ag_write_unit_files(ctx, tmp_dir, u);
                lap("write unit files");
            }
            std::lock_guard<std::mutex> lk(mu);
            if (rc && !first_rc) first_rc = rc;
            if (done) done(u, rc, rc ? ctx->err.c_str() : "", user);
            if (rc) { stop = true; cv.notify_all(); }
        }
    };
    std::vector<std::thread> th;
    auto load_reads = [&]() -> int {
        long long wlo = -1, whi = -1;
        if (ctxs[0]->reads_window && !host_sam && n_units) units_read_window(tmp, units, wlo, whi);
        int rc = guard(ctxs[0], [&] { load_reads_impl(ctxs[0], reads_fa, wlo, whi); });
        if (!rc && n_ctx > 1) rc = ag_broadcast_reads(ctxs, n_ctx, nullptr);
        return rc;
    };
    int rc_reads = 0;
    if (reads_fa && host_sam) rc_reads = load_reads();                       // the host SAM parser needs the read lengths: reads first
    if (!rc_reads) for (int i = 0; i < n_prep && n_units; i++) th.emplace_back(preparer);
    lap("start preparers");
    if (reads_fa && !host_sam) rc_reads = load_reads();                      // raw text -> GPU while the preparers parse the first units' genome / PSL
    lap("reads text -> device");
    if (rc_reads) { stop = true; cv.notify_all(); for (auto& t : th) t.join(); return rc_reads; }
    for (int i = 1; i < n_ctx; i++) th.emplace_back(worker, ctxs[i]);
    worker(ctxs[0]);
    for (auto& t : th) t.join();
    return first_rc;
}

int ag_run_units_files(ag_ctx** ctxs, int n_ctx, const char* tmp_dir, int first_unit, int n_units, int prefetch,
                       void (*done)(int unit, int rc, const char* error, void* user), void* user) {
    if (n_units < 0) return 1;
    std::vector<int> units; for (int u = 0; u < n_units; u++) units.push_back(first_unit + u);
    return run_units_impl(ctxs, n_ctx, tmp_dir, units, prefetch, nullptr, done, user);
}
int ag_run_job_files(ag_ctx** ctxs, int n_ctx, const char* tmp_dir, const char* reads_fa, const int* units, int n_units, int prefetch,
                     void (*done)(int unit, int rc, const char* error, void* user), void* user) {
    if (n_units < 0 || (n_units && !units)) return 1;
    return run_units_impl(ctxs, n_ctx, tmp_dir, std::vector<int>(units, units + n_units), prefetch, reads_fa, done, user);
}

int ag_get_unit(ag_ctx* ctx, ag_unit_view* out) {
    return guard(ctx, [&] {
        AgUnit& u = ctx->unit;
        if (!u.cdesc.empty() && u.chain_pos.empty()) { ctx->dev->unpin_all(); ag_expand_chains(u); ctx->uploaded = false; }   // inspection copy of the chain-major arrays
        if (u.aln_on_device && u.aln.empty()) { ctx->dev->unpin_all(); ctx->dev->fetch_alignments(u.aln, u.ext); }   // inspection copy; the device keeps using its own
        if (u.cm_start.size() != u.ref.size() + 1 || u.cm.size() != u.chain_pos.size()) { ctx->dev->unpin_all(); ag_expand_contimers(u); }   // table view on demand
        out->ref = u.ref.data(); out->n_ref = u.n_ref; out->n_tail = (uint32_t)u.ref.size() - u.n_ref;
        out->cm_start = u.cm_start.data(); out->cm = (const ag_cm_c*)u.cm.data(); out->n_cm = (uint32_t)u.cm.size();
        out->threads = (const ag_cthread_c*)u.threads.data(); out->n_threads = (uint32_t)u.threads.size();
        out->chain_pos = u.chain_pos.data(); out->chain_base = u.chain_base.data();
        out->aln = (const ag_aln_c*)u.aln.data(); out->n_aln = u.aln.size(); out->ext = (const ag_seg_c*)u.ext.data(); out->n_ext = u.ext.size();
    });
}

int ag_get_stats(ag_ctx* ctx, ag_stats* o) {
    return guard(ctx, [&] {
        const AgTimings& t = ctx->dev->timings();
        memset(o, 0, sizeof *o);
        o->ms_h2d = t.h2d; o->ms_prep = t.prep; o->ms_sort = t.sort; o->ms_nodes = t.nodes; o->ms_finalize = t.finalize; o->ms_edges = t.edges;
        o->ms_components = t.components; o->ms_chains = t.chains; o->ms_walk = t.walk; o->ms_materialize = t.materialize; o->ms_d2h = t.d2h;
        o->s_parse = ctx->s_parse; o->s_device_section = ctx->s_device; o->s_post = ctx->s_post;
        o->n_aln = ctx->n_aln; o->n_nodes = t.n_nodes; o->n_walks = ctx->n_walks; o->n_emitted = ctx->n_emitted; o->n_keys = t.n_keys; o->n_tiles = t.n_tiles;
        o->ms_ingest_reads = t.ingest_reads; o->ms_ingest_sam = t.ingest_sam; o->sam_device = t.sam_device; o->sam_host = t.sam_host; o->reads_device = t.reads_device; o->reads_host = t.reads_host; o->regrows = (uint64_t)t.regrows; o->reads_windowed = t.reads_windowed; o->ms_stage = t.stage; o->ms_build_kernel = t.build_kernel; o->ms_select = t.select;
        o->kernel_launches = ctx->dev->kernel_launches(); o->h2d_bytes = t.h2d_bytes; o->d2h_bytes = t.d2h_bytes; o->walk_fallback = t.walk_fallback;
    });
}
int ag_reset_stats(ag_ctx* ctx) {
    return guard(ctx, [&] { ctx->dev->reset_timings(); ctx->s_parse = ctx->s_device = ctx->s_post = 0; ctx->n_aln = ctx->n_walks = ctx->n_emitted = 0; });
}
int ag_keep_node_counts(ag_ctx* ctx, int on) {
    return guard(ctx, [&] { ctx->dev->set_keep_counts(on != 0); });
}

int ag_dump_nodes_text(ag_ctx* ctx, const char** text, uint64_t* len) {
    return guard(ctx, [&] {
        ensure_host_reads(ctx);
        AgNodeDump d; ctx->dev->dump_nodes(d);
        ag_format_node_dump(d, (*ctx->rp), ctx->dump);
        *text = ctx->dump.data(); *len = ctx->dump.size();
    });
}
void* ag_cuda_stream(ag_ctx* ctx) { return ctx ? ctx->dev->stream() : nullptr; }
int ag_invalidate_device_inputs(ag_ctx* ctx) { return guard(ctx, [&] { ctx->uploaded = false; }); }
int ag_reupload_reads(ag_ctx* ctx) {
    return guard(ctx, [&] {
        if (!ctx->have_reads) throw AgHostError{"reads not set"};
        ctx->reads_dirty = true;
    });
}
int ag_formalize_inputs(ag_ctx* ctx, const char* contig_fa, const char* genome_fa, const char* tmp_dir, int part, int* n_units) {
    return guard(ctx, [&] {
        std::vector<std::string> cids, gids;
        ag_formalize_contigs(contig_fa, tmp_dir, cids);
        *n_units = ag_formalize_genome(genome_fa, tmp_dir, part, gids);
    });
}
int ag_pin_staged(ag_ctx* ctx) {
    return guard(ctx, [&] {
        ensure_host_reads(ctx);
        AgDevice& d = *ctx->dev; const AgReads& r = (*ctx->rp); const AgUnit& u = ctx->unit;
        d.unpin_all();
        d.pin(r.bases.data(), r.bases.size() * 4); d.pin(r.nmask.data(), r.nmask.size() * 4); d.pin(r.len.data(), r.len.size() * 2);
        d.pin(u.ref.data(), u.ref.size()); if (u.threads.empty()) { d.pin(u.cm_start.data(), u.cm_start.size() * 4); d.pin(u.cm.data(), u.cm.size() * sizeof(ag_cm)); }
        d.pin(u.threads.data(), u.threads.size() * sizeof(ag_cthread));
        d.pin(u.chain_pos.data(), u.chain_pos.size() * 4); d.pin(u.chain_base.data(), u.chain_base.size());
        d.pin(u.aln.data(), u.aln.size() * sizeof(ag_aln)); d.pin(u.ext.data(), u.ext.size() * sizeof(ag_seg));
    });
}
// the aligner call of refinement() (AG:2957-2983) without an aligner: seed-and-verify containment search, candidate verification on the GPU
int ag_containment_search_files(ag_ctx* ctx, const char* db_fa, const char* query_fa, const char* out_psl) {
    return guard(ctx, [&] {
        ag_contain_search(db_fa, query_fa, out_psl,
            [](const AgSeqSet& db, const AgSeqSet& qs, const std::vector<AgPlacement>& cand, std::vector<u32>& match, void* p) { ((ag_ctx*)p)->dev->verify_placements(db, qs, cand, match); }, ctx);
    });
}
// removeMisassembly (AG:4281-4297) for one output file; the per-base coverage pile-up runs on the context's GPU
int ag_remove_misassembly_file(ag_ctx* ctx, const char* file, const char* id, int coverage, const char* tmp_dir, int (*align)(const char* id, void* user), void* user) {
    return guard(ctx, [&] {
        struct U { ag_ctx* ctx; int (*align)(const char*, void*); void* user; } u{ctx, align, user};
        ag_remove_misassembly(file, id, coverage, tmp_dir,
            [](const std::string& id_, void* p) -> bool { U* q = (U*)p; return q->align ? q->align(id_.c_str(), q->user) != 0 : true; },
            [](const std::string& sam, const std::vector<u32>& len, std::vector<int>& cov, void* p) {
                U* q = (U*)p;
                if (!q->ctx->host_parse && q->ctx->dev->coverage_pileup(sam, len, cov)) return;
                ag_coverage_pileup_host(sam, len, cov, nullptr);
            }, &u);
    });
}
int ag_set_option(ag_ctx* ctx, const char* name, long value) {
    return guard(ctx, [&] {
        const std::string n = name ? name : "";
        if (n == "host_parse") ctx->host_parse = value != 0;
        else if (n == "reads_window") ctx->reads_window = value != 0;
        else if (n == "node_cap" || n == "ovf_cap" || n == "eovf_cap" || n == "key_cap" || n == "cand_cap" || n == "hwalk_cap" || n == "rank_rounds" || n == "tma" || n == "bases_cap" || n == "fused_extend" || n == "scan_onepass") ctx->dev->set_option(n, value);
        else throw AgHostError{"unknown option: " + n};
    });
}
int ag_timer_start(ag_ctx* ctx) { return guard(ctx, [&] { ctx->dev->timer_start(); }); }
int ag_timer_stop(ag_ctx* ctx, float* ms) { return guard(ctx, [&] { *ms = ctx->dev->timer_stop(); }); }

}  // extern "C"

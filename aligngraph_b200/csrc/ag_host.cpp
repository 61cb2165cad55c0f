// Host side of the B200 AlignGraph hot path — see ag_host.h.  Reference citations: AG:line = AlignGraph/AlignGraph.cpp.
#include <sched.h>
#include <condition_variable>
#include "ag_host.h"
#include "ag_samcore.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <climits>
#include <cerrno>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <malloc.h>
#include <sys/resource.h>
#include <thread>
#include <chrono>
#include <mutex>
#include <atomic>
#include <memory>

namespace {

const int kBatchPairs = 1000000;  // BATCH, AG:37
const double kReadThreshold = 0.6, kContigThreshold = 0.5, kInitContigThreshold = 0.5;  // AG:28-34
const long kLargeChunk = 1000000;  // AG:40

struct FileMap {
    const char* p = nullptr; size_t n = 0; int fd = -1; bool ok = false;
    explicit FileMap(const std::string& path) {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) != 0) return;
        n = (size_t)st.st_size; ok = true;
        if (n) { void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0); if (m == MAP_FAILED) { ok = false; n = 0; } else { p = (const char*)m; madvise(m, n, MADV_SEQUENTIAL); } }
    }
    ~FileMap() { if (p) munmap((void*)p, n); if (fd >= 0) close(fd); }
};

// The reference reads every file with `while(in.good()) { getline(in, buf); if(buf[0] == 0) break; ... }`.  next() hands out
// exactly the lines that loop sees: an empty line (or the end of a file that ends in '\n') yields n == 0 — the `break`.
struct Lines {
    const char* p; const char* end; bool good = true;
    Lines(const char* b, size_t n) : p(b), end(b + n) {}
    bool next(const char*& s, size_t& n) {
        if (!good) return false;
        if (p == end) { good = false; s = p; n = 0; return true; }
        const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
        s = p;
        if (e) { n = (size_t)(e - p); p = e + 1; } else { n = (size_t)(end - p); p = end; good = false; }
        return true;
    }
};

int ag_atoi(const char* s, size_t n) {  // atoi() over a bounded field
    size_t i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    unsigned long long v = 0; bool sat = false;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) { v = v * 10 + (unsigned)(s[i] - '0'); if (v > (unsigned long long)LONG_MAX) sat = true; }
    long r = sat ? (neg ? LONG_MIN : LONG_MAX) : (neg ? -(long)v : (long)v);
    return (int)r;
}

struct Out {  // buffered text sink: a file, or an in-memory string (written in place, no intermediate copy)
    FILE* f = nullptr; std::string own; std::string* b;
    explicit Out(const std::string& path) : f(fopen(path.c_str(), "wb")), b(&own) { if (!f) throw AgHostError{"CANNOT OPEN FILE!"}; own.reserve(1 << 21); }
    explicit Out(std::string* m) : b(m) { m->clear(); }
    ~Out() { if (f) { flush(); fclose(f); } }
    void flush() { if (f && !b->empty()) { fwrite(b->data(), 1, b->size(), f); b->clear(); } }
    void maybe_flush() { if (f && b->size() > (1u << 20)) flush(); }
    void put(const char* s, size_t n) { b->append(s, n); maybe_flush(); }
    void put(const std::string& s) { put(s.data(), s.size()); }
    void ch(char c) { b->push_back(c); }
    void num(unsigned long v) { char t[24]; int i = 24; do { t[--i] = (char)('0' + v % 10); v /= 10; } while (v); b->append(t + i, (size_t)(24 - i)); }
    void inum(long v) { if (v < 0) { ch('-'); num((unsigned long)(-(v + 1)) + 1); } else num((unsigned long)v); }
    void wrap60(const char* s, size_t n) {  // 60 columns, newline after the last base (AG:2179-2184)
        size_t old = b->size(), lines = (n + 59) / 60;
        b->resize(old + n + lines);
        char* p = &(*b)[old];
        for (size_t i = 0; i < n; i += 60) { size_t m = std::min<size_t>(60, n - i); memcpy(p, s + i, m); p += m; *p++ = '\n'; }
        maybe_flush();
    }
    void wrap60(const std::string& s) { wrap60(s.data(), s.size()); }
};

inline int absdiff(u32 a, u32 b) { int d = (int)(a - b); return d < 0 ? -d : d; }

void revcomp(std::string& s) {  // AG:854-865
    std::reverse(s.begin(), s.end());
    for (char& c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

}  // namespace

// =============================================================================================================================
// reads
// =============================================================================================================================
static inline int base_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
struct BaseLut { unsigned char v[256]; BaseLut() { for (int i = 0; i < 256; i++) v[i] = 4; v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } unsigned char operator[](unsigned char c) const { return v[c]; } };
static const BaseLut kBaseLut;   // 0-3 for ACGT, 4 for anything else (AG:1349 counts those as 'N'; lower case is not ACGT for the reference)

static void pack_init(AgReads& out, size_t n_reads, u32 maxlen) {
    out.stride2 = (maxlen + 15) / 16; out.stridem = (maxlen + 31) / 32;
    if (!out.stride2) out.stride2 = out.stridem = 1;
    out.bases.assign_zero(n_reads * out.stride2); out.nmask.assign_zero(n_reads * out.stridem);
    out.n_pairs = n_reads / 2; out.len.assign(out.n_pairs, 0); out.exc.clear(); out.exc_complete = true;   // every packer below records each mask bit in exc
}
static void pack_one(AgReads& out, size_t r, const char* s, size_t n) {
    u32* b = &out.bases[r * out.stride2]; u32* m = &out.nmask[r * out.stridem];
    for (size_t i = 0; i < n; i++) {
        int c = base_code(s[i]);
        if (c < 0) { m[i >> 5] |= 1u << (i & 31); out.exc.push_back({(u64)r * 65536 + i, s[i]}); }
        else b[i >> 4] |= (u32)c << ((i & 15) * 2);
    }
}

void ag_pack_reads(const std::vector<std::string>& seqs, AgReads& out) {
    size_t maxlen = 0;
    for (auto& s : seqs) maxlen = std::max(maxlen, s.size());
    if (maxlen > 65535) throw AgHostError{"READ TOO LONG"};
    pack_init(out, seqs.size() & ~(size_t)1, (u32)maxlen);
    for (size_t r = 0; r + 1 < seqs.size(); r += 2) {
        if (seqs[r].size() != seqs[r + 1].size()) throw AgHostError{"INCONSISTENT PE FILES!"};
        out.len[r / 2] = (uint16_t)seqs[r].size();
        pack_one(out, r, seqs[r].data(), seqs[r].size()); pack_one(out, r + 1, seqs[r + 1].data(), seqs[r + 1].size());
    }
}

static size_t parallel_min_bytes() {  // files smaller than this are parsed sequentially (AG_PARSE_PARALLEL_MIN overrides, for tests)
    if (const char* e = getenv("AG_PARSE_PARALLEL_MIN")) return (size_t)atol(e);
    return (size_t)1 << 22;
}
static thread_local int t_thread_budget = 0;
void ag_set_thread_budget(int n) { t_thread_budget = n; }
static int host_threads() {
    if (t_thread_budget > 0) return t_thread_budget;   // a caller that runs several parsers side by side shares the cores out (ag_run_units_files)
    if (const char* e = getenv("AG_THREADS")) { int n = atoi(e); if (n > 0) return n; }
    unsigned h = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(h ? h : 1, 32);
}

// ---- host thread team: persistent workers for the text formatters (the reference is single-threaded; these loops are pure copies) ----
namespace {
struct Team {
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, cv_done;
    const std::function<void(int)>* job = nullptr;
    int n_chunks = 0, active = 0; unsigned long gen = 0; bool stop = false;
    std::atomic<int> next{0};
    std::mutex run_m;   // one job at a time (several contexts may share the team)
    int size = 1;
    Team() {
        int T = 0;
        if (const char* e = getenv("AG_THREADS")) T = atoi(e);
        if (T <= 0) {
            cpu_set_t set; CPU_ZERO(&set);
            T = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            T = std::min(T, 32);
        }
        size = std::max(T, 1);
        for (int i = 1; i < size; i++) th.emplace_back([this] { work(); });
    }
    ~Team() { { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
    void drain() { for (;;) { int c = next.fetch_add(1); if (c >= n_chunks) break; (*job)(c); } }
    void work() {
        unsigned long seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return stop || gen != seen; }); if (stop) return; seen = gen; }
            drain();
            { std::lock_guard<std::mutex> l(m); if (--active == 0) cv_done.notify_one(); }
        }
    }
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (size == 1 || n == 1) { for (int c = 0; c < n; c++) fn(c); return; }
        std::lock_guard<std::mutex> rl(run_m);
        { std::lock_guard<std::mutex> l(m); job = &fn; n_chunks = n; next = 0; active = (int)th.size(); gen++; }
        cv.notify_all();
        drain();
        std::unique_lock<std::mutex> l(m); cv_done.wait(l, [&] { return active == 0; });
    }
};
Team& team() { static Team t; return t; }
}  // namespace
void ag_parallel_chunks(int n_chunks, const std::function<void(int)>& fn) { team().run(n_chunks, fn); }
int ag_team_size() { return team().size; }

// Fast path for the file formalizeInput writes (AG:3455-3471): strictly alternating ">id" / one sequence line.  Chunks of the mapped
// file are scanned and packed by several threads; anything irregular (multi-line record, empty line, odd record count, unequal mates)
// returns false and the sequential parser below — the literal semantics — takes over.
static bool parse_reads_parallel(const char* p, size_t n, AgReads& out) {
    const int T = host_threads();
    if (T <= 1 || n < parallel_min_bytes()) return false;
    if (n == 0 || p[0] != '>' || p[n - 1] != '\n') return false;  // (empty lines are caught by the parallel scan below)
    std::vector<size_t> cut((size_t)T + 1, n);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {  // chunk starts: the next '>' that begins a line
        size_t o = n / T * t;
        const char* q = p + o;
        for (;;) { q = (const char*)memchr(q, '>', (size_t)(p + n - q)); if (!q) break; if (q[-1] == '\n') break; q++; }
        cut[t] = q ? (size_t)(q - p) : n;
    }
    for (int t = 1; t <= T; t++) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<size_t> nrec((size_t)T, 0), maxlen((size_t)T, 0);
    std::vector<char> bad((size_t)T, 0);
    auto scan = [&](int t) {
        const char* q = p + cut[t]; const char* e = p + cut[t + 1];
        size_t cnt = 0, mx = 0;
        while (q < e) {
            if (*q != '>') { bad[t] = 1; return; }
            const char* h = (const char*)memchr(q, '\n', (size_t)(e - q));
            if (!h || h + 1 >= e) { bad[t] = 1; return; }
            const char* sq = h + 1;
            if (*sq == '>') { bad[t] = 1; return; }
            const char* l = (const char*)memchr(sq, '\n', (size_t)(e - sq));
            if (!l) { bad[t] = 1; return; }
            if (l == sq) { bad[t] = 1; return; }   // an empty line ends the file for the reference (AG:375): leave it to the sequential parser
            mx = std::max(mx, (size_t)(l - sq)); cnt++;
            q = l + 1;
        }
        nrec[t] = cnt; maxlen[t] = mx;
    };
    auto tr0 = std::chrono::steady_clock::now();
    { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(scan, t); for (auto& x : th) x.join(); }
    auto tr1 = std::chrono::steady_clock::now();
    size_t total = 0, mx = 0;
    std::vector<size_t> base((size_t)T + 1, 0);
    for (int t = 0; t < T; t++) { if (bad[t]) return false; base[t] = total; total += nrec[t]; mx = std::max(mx, maxlen[t]); }
    if (total & 1 || mx > 65535) return false;
    pack_init(out, total, (u32)mx);
    std::vector<std::vector<std::pair<u64, char>>> exc((size_t)T);
    std::vector<uint16_t> rlen(total);
    auto pack = [&](int t) {
        const char* q = p + cut[t]; const char* e = p + cut[t + 1];
        size_t r = base[t];
        u32* B = out.bases.data(); u32* M = out.nmask.data();
        while (q < e) {
            const char* sq = (const char*)memchr(q, '\n', (size_t)(e - q)) + 1;
            const char* l = (const char*)memchr(sq, '\n', (size_t)(e - sq));
            size_t len = (size_t)(l - sq);
            u32* b = B + r * out.stride2; u32* m = M + r * out.stridem;
            for (size_t i = 0; i < len; i += 16) {   // one packed word at a time, no read-modify-write per base
                const size_t w16 = std::min<size_t>(16, len - i);
                u32 w = 0, nonacgt = 0;
                size_t j = 0;
                for (; j + 8 <= w16; j += 8) {   // eight characters per step in a 64-bit word
                    uint64_t x; memcpy(&x, sq + i + j, 8);
                    const uint64_t c = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ull;             // A C G T -> 0 1 2 3 in every byte
                    const uint64_t c0 = c & 0x0101010101010101ull, c1 = (c >> 1) & 0x0101010101010101ull;
                    const uint64_t expect = 0x4141414141414141ull + c0 * 2 + c1 * 6 + (c0 & c1) * 0x0B;   // 'A' 'C' 'G' 'T' rebuilt from the code
                    if (x == expect) {
                        uint64_t y = (c | (c >> 6)) & 0x000F000F000F000Full;
                        y = (y | (y >> 12)) & 0x000000FF000000FFull;
                        y = (y | (y >> 24)) & 0xFFFFull;
                        w |= (u32)y << (2 * j);
                    } else {
                        for (size_t t8 = 0; t8 < 8; t8++) { const u32 cc = kBaseLut[(unsigned char)sq[i + j + t8]]; w |= (cc & 3u) << (2 * (j + t8)); nonacgt |= (cc >> 2) << (j + t8); }
                    }
                }
                for (; j < w16; j++) { const u32 cc = kBaseLut[(unsigned char)sq[i + j]]; w |= (cc & 3u) << (2 * j); nonacgt |= (cc >> 2) << j; }
                b[i >> 4] = w;
                if (nonacgt) {
                    m[i >> 5] |= nonacgt << (i & 31);
                    for (size_t t16 = 0; t16 < w16; t16++) if ((nonacgt >> t16) & 1) exc[t].push_back({(u64)r * 65536 + i + t16, sq[i + t16]});
                }
            }
            rlen[r] = (uint16_t)len;
            r++; q = l + 1;
        }
    };
    auto tr2 = std::chrono::steady_clock::now();
    { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(pack, t); for (auto& x : th) x.join(); }
    auto tr3 = std::chrono::steady_clock::now();
    for (size_t i = 0; i + 1 < total; i += 2) { if (rlen[i] != rlen[i + 1]) throw AgHostError{"INCONSISTENT PE FILES!"}; out.len[i / 2] = rlen[i]; }
    for (int t = 0; t < T; t++) out.exc.insert(out.exc.end(), exc[t].begin(), exc[t].end());
    if (getenv("AG_POST_TIMING")) { auto tr4 = std::chrono::steady_clock::now(); auto ms = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count() * 1e3; };
        fprintf(stderr, "  [reads] scan %.0f ms, alloc %.0f ms, pack %.0f ms, tail %.0f ms (%d threads)\n", ms(tr0, tr1), ms(tr1, tr2), ms(tr2, tr3), ms(tr3, tr4), T); }
    return true;
}

void ag_parse_reads(const std::string& path, AgReads& out) {
    FileMap fm(path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    if (parse_reads_parallel(fm.p, fm.n, out)) return;
    // pass 1: record boundaries (a record = '>' line + the sequence lines up to the next '>'), AG:372-396
    struct Rec { const char* s; size_t n; bool multi; };
    std::vector<Rec> recs;
    std::vector<std::string> joined;  // only for multi-line records
    {
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') recs.push_back(Rec{nullptr, 0, false});
            else if (!recs.empty()) {
                Rec& r = recs.back();
                if (!r.s) { r.s = s; r.n = n; }
                else {  // sequence spread over several lines
                    if (!r.multi) { joined.emplace_back(r.s, r.n); r.multi = true; r.n = joined.size() - 1; }
                    joined[r.n].append(s, n);
                }
            }
        }
    }
    size_t maxlen = 0;
    for (auto& r : recs) maxlen = std::max(maxlen, r.multi ? joined[r.n].size() : r.n);
    if (maxlen > 65535) throw AgHostError{"READ TOO LONG"};
    pack_init(out, recs.size() & ~(size_t)1, (u32)maxlen);
    for (size_t r = 0; r + 1 < recs.size(); r += 2) {
        const char* s1 = recs[r].multi ? joined[recs[r].n].data() : recs[r].s; size_t n1 = recs[r].multi ? joined[recs[r].n].size() : recs[r].n;
        const char* s2 = recs[r + 1].multi ? joined[recs[r + 1].n].data() : recs[r + 1].s; size_t n2 = recs[r + 1].multi ? joined[recs[r + 1].n].size() : recs[r + 1].n;
        if (n1 != n2) throw AgHostError{"INCONSISTENT PE FILES!"};
        out.len[r / 2] = (uint16_t)n1;
        pack_one(out, r, s1, n1); pack_one(out, r + 1, s2, n2);
    }
}

char AgReads::at(u32 read, u32 rc, u32 rlen, u32 off) const {
    u32 i = rc ? rlen - 1 - off : off;
    if (device_only()) {
        if (exc.empty()) return 0;
        u64 key = (u64)read * 65536 + i;
        auto it = std::lower_bound(exc.begin(), exc.end(), std::make_pair(key, (char)CHAR_MIN));
        return it != exc.end() && it->first == key ? it->second : 0;
    }
    if ((nmask[(size_t)read * stridem + (i >> 5)] >> (i & 31)) & 1) {
        u64 key = (u64)read * 65536 + i;
        auto it = std::lower_bound(exc.begin(), exc.end(), std::make_pair(key, (char)CHAR_MIN));
        return it != exc.end() && it->first == key ? it->second : 'N';
    }
    u32 c = (bases[(size_t)read * stride2 + (i >> 4)] >> ((i & 15) * 2)) & 3;
    return "ACGT"[rc ? 3 - c : c];
}

// =============================================================================================================================
// unit genome
// =============================================================================================================================
void ag_load_genome(const std::string& path, AgUnit& u) {
    FileMap fm(path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    u.ref.clear();
    Lines ln(fm.p, fm.n); const char* s; size_t n;
    while (ln.next(s, n)) {
        if (n == 0 || s[0] == 0) break;
        if (s[0] == '>') continue;
        u.ref.append(s, n);
    }
    u.n_ref = (u32)u.ref.size();
}

// =============================================================================================================================
// contigs: PSL -> position sets -> contiMer threads
// =============================================================================================================================
namespace {
struct Chunk { std::string bases; int id = 0; std::vector<std::vector<u32>> sets; std::vector<int> fr; int outputted = 0; };
struct PslRec { u32 tid, tstart, tend, tgap, sid, sstart, send, sgap, ssize, fr; std::vector<ag_seg> segs; };

void parse_psl(const char* s, size_t n, PslRec& r) {  // AG:406-522, fields by tab index
    const char* f[21]; size_t fl[21]; int nf = 0;
    const char* p = s; const char* end = s + n;
    while (nf < 21) {
        const char* e = (const char*)memchr(p, '\t', (size_t)(end - p));
        f[nf] = p; fl[nf] = e ? (size_t)(e - p) : (size_t)(end - p); nf++;
        if (!e) break;
        p = e + 1;
    }
    auto fi = [&](int i) { return i < nf ? ag_atoi(f[i], fl[i]) : 0; };
    r.tid = (u32)fi(13); r.tstart = (u32)fi(15); r.tend = (u32)fi(16); r.tgap = (u32)fi(7);
    r.sstart = (u32)fi(11); r.send = (u32)fi(12); r.sgap = (u32)fi(5); r.ssize = (u32)fi(10);
    r.fr = AG_NONE;
    if (nf > 8 && fl[8] > 0) r.fr = f[8][0] == '+' ? 0u : 1u;
    r.sid = 0;
    if (nf > 9) { const char* dot = (const char*)memchr(f[9], '.', fl[9]); r.sid = (u32)ag_atoi(f[9], dot ? (size_t)(dot - f[9]) : fl[9]); }
    r.segs.clear();
    auto list = [&](int i, int which) {
        if (i >= nf) return;
        const char* q = f[i]; const char* qe = f[i] + fl[i]; size_t k = 0;
        for (;;) {
            const char* c = (const char*)memchr(q, ',', (size_t)(qe - q));
            if (!c) break;
            u32 v = (u32)ag_atoi(q, (size_t)(c - q));
            if (which == 0) { ag_seg sg; sg.src = sg.dst = AG_NONE; sg.len = v; r.segs.push_back(sg); }
            else if (k < r.segs.size()) { if (which == 1) r.segs[k].src = v; else r.segs[k].dst = v; }
            k++; q = c + 1;
        }
    };
    list(18, 0); list(19, 1); list(20, 2);
}

// keepPositions (AG:731-748): is the most recent position set of chunk `sid` at least `thr` aligned?
int keep_set(std::vector<Chunk>& ch, u32 sid, double thr) {
    if (sid == AG_NONE) return 1;
    if (ch[sid].sets.empty()) return 1;
    const std::vector<u32>& last = ch[sid].sets.back();
    int match = 0;
    for (u32 v : last) if (v != AG_NONE) match++;
    return (double)match / last.size() >= thr ? 1 : 0;
}
}  // namespace

// tmp/_contigs.fa is the same for every unit (the reference re-reads it per chromosome, AG:1225-1228): parse it once per file version
namespace {
struct ChunkCache { std::string path; long size = -1, mtime = -1, mtime_ns = -1; std::shared_ptr<const AgChunkStore> store; };
std::mutex g_chunk_mu;
ChunkCache g_chunk_cache;
std::atomic<u64> g_chunk_version{1};
}
std::shared_ptr<const AgChunkStore> ag_chunk_store(const std::string& contigs_fa) {
    std::lock_guard<std::mutex> lk(g_chunk_mu);
    struct stat st;
    if (stat(contigs_fa.c_str(), &st) != 0) throw AgHostError{"CANNOT OPEN FILE!"};
    ChunkCache& c = g_chunk_cache;
    if (!c.store || c.path != contigs_fa || c.size != (long)st.st_size || c.mtime != (long)st.st_mtim.tv_sec || c.mtime_ns != (long)st.st_mtim.tv_nsec) {
        FileMap fm(contigs_fa);
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        auto sp = std::make_shared<AgChunkStore>();
        sp->blob.reserve(fm.n);
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {   // AG:322-359
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') {
                const char* dot = (const char*)memchr(s, '.', n);
                sp->id.push_back(dot ? ag_atoi(dot + 1, (size_t)(s + n - dot - 1)) : 0);
                sp->off.push_back(sp->blob.size());
            } else if (!sp->id.empty()) sp->blob.append(s, n);
        }
        sp->off.push_back(sp->blob.size());
        sp->version = g_chunk_version.fetch_add(1);
        c.store = sp; c.path = contigs_fa; c.size = (long)st.st_size; c.mtime = (long)st.st_mtim.tv_sec; c.mtime_ns = (long)st.st_mtim.tv_nsec;
    }
    return c.store;
}
static void load_chunks(const std::string& contigs_fa, std::vector<Chunk>& ch) {
    std::shared_ptr<const AgChunkStore> st = ag_chunk_store(contigs_fa);
    ch.resize(st->n());
    for (size_t i = 0; i < ch.size(); i++) { ch[i].id = st->id[i]; ch[i].bases.assign(st->blob, (size_t)st->off[i], st->size(i)); }
}

void ag_expand_contimers(const ag_cthread* threads, size_t n_threads, const u32* chain_pos, size_t n_cm, size_t n_pos, std::vector<u32>& cm_start, std::vector<ag_cm>& cm) {
    cm_start.assign(n_pos + 1, 0);
    for (size_t k = 0; k < n_cm; k++) cm_start[chain_pos[k] + 1]++;
    for (size_t i = 0; i < n_pos; i++) cm_start[i + 1] += cm_start[i];
    cm.assign(n_cm, ag_cm{});
    std::vector<u32> fill(cm_start.begin(), cm_start.end() - 1);
    for (size_t i = 0; i < n_threads; i++) {
        const ag_cthread& t = threads[i];
        for (u32 k = t.first; k <= t.term; k++) {   // threads are stored in push order, so every position's list comes out in push order
            ag_cm m; m.cid = t.cid; m.coff = k < t.term ? t.coff_first + (k - t.first) : t.coff_term; m.chain = k; m.term = t.term;
            cm[fill[chain_pos[k]]++] = m;
        }
    }
}
void ag_expand_contimers(AgUnit& u) { ag_expand_contimers(u.threads.data(), u.threads.size(), u.chain_pos.data(), u.chain_pos.size(), u.ref.size(), u.cm_start, u.cm); }

void ag_thread_contigs(const std::string& contigs_fa, const std::string& psl_path, std::string& initial_text, AgUnit& u) {
    // ---- chunks (AG:322-359) ----
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (getenv("AG_POST_TIMING")) { auto t = std::chrono::steady_clock::now(); struct rusage ru; getrusage(RUSAGE_SELF, &ru); fprintf(stderr, "  [contigs] %s %.1f ms (minflt %ld, stime %.0f ms)\n", what, std::chrono::duration<double>(t - T0).count() * 1e3, ru.ru_minflt, ru.ru_stime.tv_sec * 1e3 + ru.ru_stime.tv_usec / 1e3); T0 = t; } };
    std::vector<Chunk> ch;
    load_chunks(contigs_fa, ch);
    lap("chunks");
    // ---- PSL -> position sets (AG:817-852 with updateContig AG:763-815) ----
    {
        FileMap fm(psl_path);
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        PslRec r; r.sid = AG_NONE;
        u32 last_source = AG_NONE;  // sourceIDBak (AG:762); it is -1 on entry (reset at AG:4781 / AG:1241)
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) {  // end of file: validate the last set of the LAST PARSED record's chunk (AG:830-836)
                if (r.sid != AG_NONE && r.sid < ch.size() && keep_set(ch, r.sid, kContigThreshold) == 0) ch[r.sid].sets.pop_back();
                break;
            }
            parse_psl(s, n, r);
            bool pass = (double)(r.send - r.sstart - r.sgap) / r.ssize >= kInitContigThreshold &&
                        (double)(r.tend - r.tstart - r.tgap) / (r.tend - r.tstart) >= kInitContigThreshold && r.ssize > 200;
            if (!pass) continue;
            if (r.tid == AG_NONE) continue;  // updateContig returns at once (AG:769)
            if (r.tid != 0 || r.sid >= ch.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
            Chunk& c = ch[r.sid];
            auto open_set = [&]() { c.sets.emplace_back(c.bases.size(), AG_NONE); c.fr.push_back((int)r.fr); };
            if (r.sid != last_source) {
                if (keep_set(ch, last_source, kContigThreshold) == 0) { ch[last_source].sets.pop_back(); ch[last_source].fr.pop_back(); }
                open_set();
                last_source = r.sid;
            } else {
                bool clash = false;
                for (size_t i = 0; i < r.segs.size() && !clash; i++)
                    for (u32 j = r.segs[i].src; j < r.segs[i].src + r.segs[i].len; j++) {
                        if (j >= c.bases.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                        if (c.sets.back()[j] != AG_NONE) { clash = true; break; }
                    }
                if (clash) {
                    if (keep_set(ch, r.sid, kContigThreshold) == 0) { c.sets.pop_back(); c.fr.pop_back(); }
                    open_set();
                }
            }
            std::vector<u32>& cur = c.sets.back();
            for (const ag_seg& sg : r.segs)
                for (u32 j = 0; j < sg.len; j++) {
                    if ((size_t)sg.src + j >= cur.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                    cur[sg.src + j] = sg.dst + j;
                }
        }
    }
    lap("psl+sets");
    // ---- thread every surviving set through the unit (AG:884-1177) ----
    // chain-major output, written in place: one thread after the other, each in walking order (the position-ordered table — CSR, push
    // order preserved — is derived from these arrays on the device)
    std::vector<u32>& cpos = u.chain_pos; std::string& cbase = u.chain_base;
    cpos.clear(); cbase.clear();
    std::vector<unsigned char> count(u.n_ref, 0);  // contiMers per position so far, saturating at 255 (only ">= 2" matters); grows with the tail
    { size_t cap = 0; for (const Chunk& c : ch) for (const auto& st : c.sets) cap += st.size() + 1; cpos.reserve(cap); cbase.reserve(cap); }
    u.threads.clear(); u.cm_start.clear(); u.cm.clear();
    u.ref.resize(u.n_ref);
    for (size_t sp = 0; sp < ch.size(); sp++) {
        Chunk& c = ch[sp];
        for (size_t pp = 0; pp < c.sets.size(); pp++) {
            const std::vector<u32>& ps = c.sets[pp];
            bool skip = false;
            for (size_t e = 0; e < pp && !skip; e++) if (absdiff(ps[0], c.sets[e][0]) < (int)c.bases.size()) skip = true;  // AG:902-907
            for (size_t i = 0; !skip && i + 1 < ps.size(); i++)
                if (ps[i] != AG_NONE) { if (ps[i] >= count.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"}; if (count[ps[i]] >= 2) skip = true; }  // AG:908-920
            if (skip) continue;
            bool flipped = false;
            if (c.fr[pp] == 1) { revcomp(c.bases); flipped = true; }
            c.outputted = 1;
            const size_t first_push = cpos.size();
            u32 coff_first = AG_NONE, coff_prev = AG_NONE, coff_last = 0; bool consecutive = true;
            auto push = [&](u32 pos, u32 coff, char base, bool terminal) {
                if (pos >= count.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                if (coff_first == AG_NONE) coff_first = coff;
                else if (!terminal && coff != coff_prev + 1) consecutive = false;   // contig offsets advance by one per contiMer up to the terminal
                coff_prev = coff; coff_last = coff;
                cpos.push_back(pos); cbase.push_back(base);
                if (count[pos] != 255) count[pos]++;
            };
            u32 cur = AG_NONE, nxt = AG_NONE;
            bool have_next = false;
            size_t i;
            for (i = 0; i + 1 < ps.size(); i++) {
                if (ps[i] == AG_NONE) continue;
                if (ps[i + 1] != AG_NONE) {
                    // a run of aligned bases i .. r-1: every one but the last has an aligned successor => ordinary steps (AG:1075-1118), appended in bulk
                    size_t r = i + 2;
                    while (r < ps.size() && ps[r] != AG_NONE) r++;
                    const size_t cnt = r - 1 - i;
                    for (size_t k = i; k < r - 1; k++) { const u32 pos = ps[k]; if (pos >= count.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"}; if (count[pos] != 255) count[pos]++; }
                    if (coff_first == AG_NONE) coff_first = (u32)i; else if ((u32)i != coff_prev + 1) consecutive = false;
                    coff_prev = coff_last = (u32)(r - 2);
                    cpos.insert(cpos.end(), ps.begin() + (long)i, ps.begin() + (long)(r - 1));
                    cbase.append(c.bases, i, cnt);
                    cur = ps[r - 2]; nxt = ps[r - 1]; have_next = true;
                    i = r - 2;   // the run's last base (index r-1) is looked at by the next iteration: it is followed by an unaligned base or is the contig's last
                    continue;
                }
                cur = ps[i]; nxt = ps[i + 1]; have_next = nxt != AG_NONE;
                char base = c.bases[i];
                if (nxt == AG_NONE) {  // bases inserted relative to the unit: append them behind the unit (SI = 0 => always "large", AG:974-1042)
                    for (size_t m = i + 2; m < ps.size(); m++) {
                        if (ps[m] == AG_NONE) continue;
                        nxt = ps[m]; have_next = true;
                        push(cur, (u32)i, base, false);
                        for (size_t j = i + 1; j < m; j++) {
                            u.ref.push_back(c.bases[j]); count.push_back(0);
                            push((u32)u.ref.size() - 1, (u32)j, c.bases[j], false);
                        }
                        i = m - 1;
                        break;
                    }
                } else push(cur, (u32)i, base, false);  // ordinary step or deletion (SD = 0 => "large", AG:1075-1118)
            }
            if (cur != AG_NONE) {
                // terminal contiMer carries the UNIT's base (AG:1121-1148)
                u32 tp = have_next ? nxt : cur;
                if (tp >= u.ref.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                push(tp, (u32)i, u.ref[tp], true);
            }
            if (cpos.size() > first_push) {
                if (!consecutive) throw AgHostError{"internal: contig thread offsets are not consecutive"};
                ag_cthread t; t.first = (u32)first_push; t.term = (u32)cpos.size() - 1; t.cid = (u32)sp; t.coff_first = coff_first; t.coff_term = coff_last;
                u.threads.push_back(t);
            }
            if (flipped) revcomp(c.bases);
        }
    }
    lap("threading");
    // ---- tmp/_initial_contigs.N.fa: original contigs with >= 50 % of their chunks threaded (AG:1179-1216) ----
    Out out(&initial_text);
    size_t c = 0, cp = 0;
    while (c < ch.size()) {
        size_t e = c; int placed = 0, total = 0; std::string whole;
        while (e < ch.size() && ch[e].id == ch[c].id) { total++; placed += ch[e].outputted; whole += ch[e].bases; e++; }
        if ((double)placed / (double)total >= kContigThreshold) { out.ch('>'); out.num(cp); out.ch('\n'); out.wrap60(whole); }
        cp++; c = e;
    }
    lap("initial_text");
}

// ---- the same in run space --------------------------------------------------------------------------------------------------------
namespace {
struct RunSet { std::vector<ag_seg> runs; u32 aligned = 0; int fr = 0; u32 ps0() const { return !runs.empty() && runs[0].src == 0 ? runs[0].dst : AG_NONE; } };
struct RChunk { std::vector<RunSet> sets; int outputted = 0; };
inline char comp_base(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
// first covered base of [a, a + n) in a sorted, disjoint run list, or NONE
u32 first_covered(const std::vector<ag_seg>& runs, u32 a, u32 n) {
    size_t lo = 0, hi = runs.size();   // first run whose end lies beyond a
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (runs[mid].src + runs[mid].len > a) hi = mid; else lo = mid + 1; }
    if (lo == runs.size() || runs[lo].src >= a + n) return AG_NONE;
    return std::max(runs[lo].src, a);
}
}  // namespace

bool ag_thread_contigs_runs(const std::string& contigs_fa, const std::string& psl_path, std::string& initial_text, AgUnit& u) {
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (getenv("AG_POST_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  [contig runs] %s %.2f ms\n", what, std::chrono::duration<double>(t - T0).count() * 1e3); T0 = t; } };
    std::shared_ptr<const AgChunkStore> store = ag_chunk_store(contigs_fa);
    const size_t nch = store->n();
    std::vector<RChunk> ch(nch);
    // ---- PSL -> position sets as run lists (loadContiAli AG:817-852 with updateContig AG:763-815 and keepPositions AG:731-748) ----
    {
        FileMap fm(psl_path);
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        PslRec r; r.sid = AG_NONE;
        u32 last_source = AG_NONE;
        auto keep = [&](u32 sid) -> int { if (sid == AG_NONE || ch[sid].sets.empty()) return 1; return (double)ch[sid].sets.back().aligned / store->size(sid) >= kContigThreshold ? 1 : 0; };
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) {
                if (r.sid != AG_NONE && r.sid < nch && keep(r.sid) == 0) {   // the reference pops the position set but not its strand flag (AG:830-836): the
                    RChunk& c = ch[r.sid];                                  // flags of the chunk's remaining sets are unaffected (nothing follows)
                    c.sets.pop_back();
                }
                break;
            }
            parse_psl(s, n, r);
            bool pass = (double)(r.send - r.sstart - r.sgap) / r.ssize >= kInitContigThreshold &&
                        (double)(r.tend - r.tstart - r.tgap) / (r.tend - r.tstart) >= kInitContigThreshold && r.ssize > 200;
            if (!pass) continue;
            if (r.tid == AG_NONE) continue;
            if (r.tid != 0 || r.sid >= nch) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
            RChunk& c = ch[r.sid];
            const u32 size = (u32)store->size(r.sid);
            auto open_set = [&]() { c.sets.emplace_back(); c.sets.back().fr = (int)r.fr; };
            if (r.sid != last_source) {
                if (keep(last_source) == 0) ch[last_source].sets.pop_back();
                open_set();
                last_source = r.sid;
            } else {
                bool clash = false;
                for (const ag_seg& sg : r.segs) {   // scan order of the reference: the first base that is out of range (error) or already placed (clash) decides
                    if (!sg.len || (u64)sg.src + sg.len > 0xFFFFFFFFull) continue;   // (a wrapped range never enters the reference's loop)
                    const u32 fc = first_covered(c.sets.back().runs, sg.src, sg.len);
                    const u64 end = (u64)sg.src + sg.len;
                    const u32 oob = end > size ? std::max(sg.src, size) : AG_NONE;
                    if (fc != AG_NONE && (oob == AG_NONE || fc < oob)) { clash = true; break; }
                    if (oob != AG_NONE) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                }
                if (clash) {
                    if (keep(r.sid) == 0) c.sets.pop_back();
                    open_set();
                }
            }
            RunSet& cur = c.sets.back();
            for (const ag_seg& sg : r.segs) {
                if (!sg.len) continue;
                if ((u64)sg.src + sg.len > size) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                if (first_covered(cur.runs, sg.src, sg.len) != AG_NONE) return false;   // a record whose own blocks overlap (later ones overwrite): per-base path
                auto it = std::lower_bound(cur.runs.begin(), cur.runs.end(), sg.src, [](const ag_seg& a, u32 v) { return a.src < v; });
                it = cur.runs.insert(it, sg);
                cur.aligned += sg.len;
                // merge with neighbours that continue the same diagonal (keeps the lists short; does not change the mapping)
                size_t k = (size_t)(it - cur.runs.begin());
                if (k + 1 < cur.runs.size() && cur.runs[k].src + cur.runs[k].len == cur.runs[k + 1].src && cur.runs[k].dst + cur.runs[k].len == cur.runs[k + 1].dst) { cur.runs[k].len += cur.runs[k + 1].len; cur.runs.erase(cur.runs.begin() + (long)k + 1); }
                if (k > 0 && cur.runs[k - 1].src + cur.runs[k - 1].len == cur.runs[k].src && cur.runs[k - 1].dst + cur.runs[k - 1].len == cur.runs[k].dst) { cur.runs[k - 1].len += cur.runs[k].len; cur.runs.erase(cur.runs.begin() + (long)k); }
            }
        }
    }
    lap("psl -> run sets");
    // ---- thread every surviving set through the unit (AG:884-1177) ----
    u.threads.clear(); u.cm_start.clear(); u.cm.clear(); u.chain_pos.clear(); u.chain_base.clear(); u.cdesc.clear(); u.cruns.clear();
    u.chunks = store;
    u.ref.resize(u.n_ref);
    std::vector<unsigned char> count(u.n_ref, 0);   // contiMers per position so far, saturating at 2 (only ">= 2" is ever asked); grows with the tail
    u32 n_chain = 0;
    for (size_t sp = 0; sp < nch; sp++) {
        RChunk& c = ch[sp];
        const u32 size = (u32)store->size(sp);
        const char* bases = store->blob.data() + store->off[sp];
        for (size_t pp = 0; pp < c.sets.size(); pp++) {
            const RunSet& rs = c.sets[pp];
            bool skip = false;
            if (size == 0) throw AgHostError{"CONTIG ALIGNMENT ERROR"};   // the reference reads ps[0] of an empty set
            for (size_t e = 0; e < pp && !skip; e++) if (absdiff(rs.ps0(), c.sets[e].ps0()) < (int)size) skip = true;   // AG:902-907
            for (size_t k = 0; !skip && k < rs.runs.size(); k++) {   // AG:908-920: any base but the last on a position that already holds two contiMers
                const ag_seg& g = rs.runs[k];
                const u32 end = std::min(g.src + g.len, size - 1);   // bases b < size - 1 only
                const u32 len = end > g.src ? end - g.src : 0;
                if (!len) continue;
                const u64 lim = std::min<u64>((u64)g.dst + len, count.size());
                const unsigned char* p = count.data(); bool two = false;
                for (u64 q = g.dst; q < lim; q++) two |= p[q] >= 2;
                if (two) { skip = true; break; }   // (a position beyond the table after a position that already holds two: the reference skips first only if
                if ((u64)g.dst + len > count.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};   //  it meets the full position first; both abort or skip the set — order kept)
            }
            if (skip) continue;
            c.outputted = 1;
            if (rs.runs.empty()) continue;
            const bool fr = rs.fr == 1;
            const u32 F = rs.runs.front().src, L = rs.runs.back().src + rs.runs.back().len - 1;
            if (F == size - 1) continue;   // only the contig's last base is aligned: the walk (i + 1 < size) never starts
            ag_cdesc d; d.first = n_chain; d.n = L - F + 1; d.F = F; d.size = size; d.fr = fr ? 1u : 0u; d.run0 = (u32)u.cruns.size(); d.nruns = (u32)rs.runs.size(); d.pad = 0; d.base_off = store->off[sp];
            for (size_t k = 0; k < rs.runs.size(); k++) {
                const ag_seg& g = rs.runs[k];
                const bool last = k + 1 == rs.runs.size();
                const u32 len = last ? g.len - 1 : g.len;   // base L is the terminal, handled below
                if ((u64)g.dst + len > count.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
                unsigned char* p = count.data() + g.dst;
                for (u32 q = 0; q < len; q++) p[q] = (unsigned char)(p[q] + (p[q] < 2));
                ag_crun cr; cr.src = g.src; cr.dst = g.dst; cr.len = g.len; cr.gap_tail = (u32)(u.ref.size() - u.n_ref);
                u.cruns.push_back(cr);
                if (!last) {   // inserted bases between this run and the next go behind the unit (AG:974-1042), one contiMer each
                    const u32 a = g.src + g.len, b = rs.runs[k + 1].src;
                    for (u32 j = a; j < b; j++) u.ref.push_back(fr ? comp_base(bases[size - 1 - j]) : bases[j]);
                    count.resize(count.size() + (b - a), 1);
                }
            }
            const u32 tp = rs.runs.back().dst + rs.runs.back().len - 1;   // terminal contiMer: the unit's base at the position of base L (AG:1121-1148)
            if (tp >= u.ref.size()) throw AgHostError{"CONTIG ALIGNMENT ERROR"};
            count[tp] = (unsigned char)(count[tp] + (count[tp] < 2));
            u.cdesc.push_back(d);
            ag_cthread t; t.first = d.first; t.term = d.first + d.n - 1; t.cid = (u32)sp; t.coff_first = F; t.coff_term = size - 1;
            u.threads.push_back(t);
            n_chain += d.n;
        }
    }
    u.n_cm_runs = n_chain;
    lap("threading");
    // ---- tmp/_initial_contigs.N.fa: original contigs with >= 50 % of their chunks threaded (AG:1179-1216); formatted by the thread team ----
    {
        struct Item { size_t c0, c1, text_off, bases; unsigned long name; };
        std::vector<Item> items;
        size_t c = 0, cp = 0, total = 0;
        while (c < nch) {
            size_t e = c; int placed = 0, tot = 0; size_t nb = 0;
            while (e < nch && store->id[e] == store->id[c]) { tot++; placed += ch[e].outputted; nb += store->size(e); e++; }
            if ((double)placed / (double)tot >= kContigThreshold) {
                size_t digits = 0; { unsigned long v = cp; do { digits++; v /= 10; } while (v); }
                Item it{c, e, total, nb, (unsigned long)cp};
                total += 1 + digits + 1 + nb + (nb + 59) / 60;
                items.push_back(it);
            }
            cp++; c = e;
        }
        initial_text.resize(total);
        char* const T = total ? &initial_text[0] : nullptr;
        ag_parallel_chunks((int)items.size(), [&](int i) {
            const Item& it = items[(size_t)i];
            char* p = T + it.text_off;
            *p++ = '>';
            { char t[24]; int k = 24; unsigned long v = it.name; do { t[--k] = (char)('0' + v % 10); v /= 10; } while (v); memcpy(p, t + k, (size_t)(24 - k)); p += 24 - k; }
            *p++ = '\n';
            // the chunks of one contig are adjacent in the blob: one 60-column wrap over the whole contig
            const char* b = store->blob.data() + store->off[it.c0];
            for (size_t o = 0; o < it.bases; o += 60) { const size_t m = std::min<size_t>(60, it.bases - o); memcpy(p, b + o, m); p += m; *p++ = '\n'; }
        });
    }
    lap("initial_text");
    return true;
}

void ag_expand_chains(AgUnit& u) {
    u.chain_pos.assign(u.n_cm_runs, 0); u.chain_base.assign(u.n_cm_runs, 'N');
    if (!u.chunks) return;
    for (const ag_cdesc& d : u.cdesc) {
        const char* bases = u.chunks->blob.data() + d.base_off;
        const ag_crun* runs = u.cruns.data() + d.run0;
        for (u32 j = 0; j < d.n; j++) {
            const u32 k = d.first + j;
            if (j == d.n - 1) { const ag_crun& g = runs[d.nruns - 1]; const u32 tp = g.dst + g.len - 1; u.chain_pos[k] = tp; u.chain_base[k] = u.ref[tp]; continue; }
            const u32 b = d.F + j;
            u32 lo = 0, hi = d.nruns;   // last run with src <= b
            while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (runs[mid].src <= b) lo = mid; else hi = mid; }
            const ag_crun& g = runs[lo];
            u.chain_pos[k] = b < g.src + g.len ? g.dst + (b - g.src) : u.n_ref + g.gap_tail + (b - g.src - g.len);
            u.chain_base[k] = d.fr ? comp_base(bases[d.size - 1 - b]) : bases[b];
        }
    }
}

// =============================================================================================================================
// SAM -> alignments
// =============================================================================================================================
namespace {
struct SamRec { u32 tid, tstart, tend, tgap, sid, sstart, send, sgap, ssize, fr; };

// AG:181-285.  Appends to `segs` (the caller clears — or, on the reference's `continue` paths, does not: AG:1258).
void parse_sam_general(const char* s, size_t n, SamRec& r, std::vector<ag_seg>& segs) {   // any number of M segments
    const char* f[6]; size_t fl[6]; int nf = 0;
    const char* p = s; const char* end = s + n;
    while (nf < 6) {
        const char* e = (const char*)memchr(p, '\t', (size_t)(end - p));
        f[nf] = p; fl[nf] = e ? (size_t)(e - p) : (size_t)(end - p); nf++;
        if (!e) break;
        p = e + 1;
    }
    for (int i = nf; i < 6; i++) { f[i] = end; fl[i] = 0; }
    r.sid = (u32)ag_atoi(f[0], fl[0]);
    r.fr = (ag_atoi(f[1], fl[1]) & 0x10) ? 1u : 0u;
    if (memchr(f[2], '*', fl[2])) { r.tid = r.tstart = r.tend = r.tgap = r.sstart = r.send = r.sgap = r.ssize = AG_NONE; return; }
    const char* dot = (const char*)memchr(f[2], '.', fl[2]);
    int pos = ag_atoi(f[3], fl[3]);
    int ins = 0, del = 0, total = 0, start = 0, end_clip = 0, lead = 1, num = 0;
    bool have_num = false;
    for (size_t i = 0; i < fl[5]; i++) {
        char c = f[5][i];
        if (c >= '0' && c <= '9') { num = have_num ? num * 10 + (c - '0') : (c - '0'); have_num = true; continue; }
        int v = have_num ? num : 0;
        if (c == 'I') { ins += v; total += v; }
        else if (c == 'D') { del += v; }
        else if (c == 'M') { ag_seg sg; sg.src = (u32)total; sg.dst = (u32)(pos + total + del - start - ins - 1); sg.len = (u32)v; segs.push_back(sg); total += v; lead = 0; }
        else if (c == 'S' && lead) { start = v; total += v; lead = 0; }
        else if (c == 'S') { end_clip = v; total += v; }
        else if (c != '*') { throw AgHostError{std::string("unknown character: ") + c}; }
        else continue;  // '*' leaves the digit buffer alone (AG:263-270)
        have_num = false; num = 0;
    }
    r.sstart = (u32)start; r.send = (u32)(total - end_clip); r.sgap = (u32)ins; r.ssize = (u32)total;
    r.tid = dot ? (u32)ag_atoi(f[2], (size_t)(dot - f[2])) : 0u;
    r.tstart = (u32)(pos - 1);
    r.tend = r.tstart + (u32)total + (u32)del - (u32)ins;
    r.tgap = (u32)del;
}

// The same through the allocation-free record parser shared with the device (ag_samcore.h); CIGARs with more M segments than it keeps
// inline go through the general version above.
void parse_sam(const char* s, size_t n, SamRec& r, std::vector<ag_seg>& segs) {
    ag_samline L;
    ag_sam_parse_line(s, (u32)n, L);
    if (L.err == AG_SAM_ERR_SEGS || n > 0xFFFFFFFFull) { parse_sam_general(s, n, r, segs); return; }
    if (L.err == AG_SAM_ERR_CHAR) throw AgHostError{std::string("unknown character: ") + L.bad_char};
    r.sid = L.sid; r.fr = L.fr; r.tid = L.tid; r.tstart = L.tstart; r.tend = L.tend; r.tgap = L.tgap;
    r.sstart = L.sstart; r.send = L.send; r.sgap = L.sgap; r.ssize = L.ssize;
    segs.insert(segs.end(), L.seg, L.seg + L.nseg);
}

// position of read offset 0 in the position set the reference would have built from `segs` (later segments overwrite)
u32 pos_at0(const std::vector<ag_seg>& segs) {
    u32 r = AG_NONE;
    for (const ag_seg& s : segs) if (s.len && s.src == 0) r = s.dst;
    return r;
}

// monotone, disjoint, merged segment list equivalent to the position set built from `segs`
void normalize(const std::vector<ag_seg>& in, u32 rlen, std::vector<ag_seg>& out) {
    out.clear();
    bool clean = true;
    for (const ag_seg& s : in) {
        if (!s.len) continue;
        if ((u64)s.src + s.len > rlen) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};  // CIGAR longer than the read: the reference writes out of bounds
        if (!out.empty()) {
            ag_seg& b = out.back();
            if (s.src < b.src + b.len || s.dst < b.dst + b.len) { clean = false; break; }
            if (s.src == b.src + b.len && s.dst == b.dst + b.len) { b.len += s.len; continue; }
        }
        out.push_back(s);
    }
    if (clean) return;
    std::vector<u32> pos(rlen, AG_NONE);
    for (const ag_seg& s : in) for (u32 j = 0; j < s.len; j++) pos[s.src + j] = s.dst + j;
    out.clear();
    for (u32 i = 0; i < rlen; i++) {
        if (pos[i] == AG_NONE) continue;
        if (!out.empty() && out.back().src + out.back().len == i && out.back().dst + out.back().len == pos[i]) { out.back().len++; continue; }
        if (!out.empty() && pos[i] < out.back().dst + out.back().len) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};  // non-monotone: unsupported
        ag_seg sg; sg.src = i; sg.dst = pos[i]; sg.len = 1; out.push_back(sg);
    }
}
}  // namespace

// self-check used by the CPU test-suite: the allocation-free record parser (ag_samcore.h) and the general one agree on one SAM line
bool ag_selfcheck_sam_line(const char* s, size_t n) {
    SamRec g; std::vector<ag_seg> gs; bool gthrow = false; char gch = 0;
    try { parse_sam_general(s, n, g, gs); } catch (const AgHostError& e) { gthrow = true; gch = e.msg.empty() ? 0 : e.msg.back(); }
    ag_samline L; ag_sam_parse_line(s, (u32)n, L);
    if (gthrow) return L.err == AG_SAM_ERR_CHAR && L.bad_char == gch;
    if (L.err == AG_SAM_ERR_CHAR) return false;
    if (g.sid != L.sid || g.fr != L.fr || g.tid != L.tid) return false;
    if (g.tid == AG_NONE) return gs.empty() && L.nseg == 0;
    if (g.tstart != L.tstart || g.tend != L.tend || g.tgap != L.tgap || g.sstart != L.sstart || g.send != L.send || g.sgap != L.sgap || g.ssize != L.ssize) return false;
    if (gs.size() != L.nseg) return false;
    if (L.nseg > AG_SAM_MAXSEG) return L.err == AG_SAM_ERR_SEGS;
    for (u32 k = 0; k < L.nseg; k++) if (gs[k].src != L.seg[k].src || gs[k].dst != L.seg[k].dst || gs[k].len != L.seg[k].len) return false;
    return true;
}
namespace {
}  // namespace

// Parallel front half of ag_parse_sam for well-formed files ('@' lines only at the top, record pairs on consecutive lines, read ids
// non-decreasing, no parse errors): threads turn text into per-pair records; the order-dependent part (1,000,000-pair batches and the
// dropped record AG:1259, duplicate rule AG:1650-1655, strand check AG:1657-1671) then runs sequentially over those records exactly
// as in the sequential parser.  Returns false (nothing changed) when the file is not of that form — the sequential parser decides.
namespace {
struct PRec { u32 sid; u32 flags; u32 p0; u32 seg_off; uint16_t n1, n2; };  // flags: bit0 fr1, bit1 fr2, bit2 passes AG:1261
}
static bool parse_sam_parallel(const char* p, size_t n, const AgReads& reads, AgUnit& u) {
    const int T = host_threads();
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (getenv("AG_POST_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  [sam] %s %.1f ms\n", what, std::chrono::duration<double>(t - T0).count() * 1e3); T0 = t; } };
    if (T <= 1 || n < parallel_min_bytes()) return false;
    size_t body = 0;
    while (body < n && p[body] == '@') { const char* l = (const char*)memchr(p + body, '\n', n - body); if (!l) return false; body = (size_t)(l - p) + 1; }
    if (body >= n || p[n - 1] != '\n' || p[body] == '\n') return false;
    // chunk boundaries on line starts; parity fixed after counting lines
    std::vector<size_t> cut((size_t)T + 1, n);
    cut[0] = body;
    for (int t = 1; t < T; t++) {
        size_t o = body + (n - body) / T * t;
        const char* l = (const char*)memchr(p + o, '\n', n - o);
        cut[t] = l ? (size_t)(l - p) + 1 : n;
    }
    for (int t = 1; t <= T; t++) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<size_t> nlines((size_t)T, 0);
    std::vector<char> empty_line((size_t)T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            size_t c = 0; const char* q = p + cut[t]; const char* e = p + cut[t + 1];
            while (q < e) { const char* l = (const char*)memchr(q, '\n', (size_t)(e - q)); if (l == q) empty_line[t] = 1; q = l + 1; c++; }
            nlines[t] = c; });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; t++) if (empty_line[t]) return false;  // an empty line ends the file for the reference (AG:1247): sequential parser
    size_t lines = 0;
    for (int t = 0; t < T; t++) {   // a chunk must start on the first line of a pair
        if (lines & 1) { const char* l = (const char*)memchr(p + cut[t], '\n', n - cut[t]); size_t nc = l ? (size_t)(l - p) + 1 : n; if (cut[t] < cut[t + 1]) { nlines[t]--; nlines[t - 1]++; lines++; } cut[t] = std::min(nc, cut[t + 1]); }
        lines += nlines[t];
    }
    if (lines & 1) return false;  // BROKEN BOWTIE FILE territory: let the sequential parser report it
    lap("line count");
    std::vector<std::vector<PRec>> recs((size_t)T);
    std::vector<std::vector<ag_seg>> segs((size_t)T);
    std::vector<char> bad((size_t)T, 0);
    auto work = [&](int t) {   // one record pair at a time through the allocation-free functions of ag_samcore.h (the per-thread code of a future kernel)
        const char* q = p + cut[t]; const char* e = p + cut[t + 1];
        recs[t].reserve((size_t)(e - q) / 140 + 16);
        segs[t].reserve((size_t)(e - q) / 70 + 16);
        ag_samline a, b;
        ag_seg n1[AG_SAM_MAXSEG], n2[AG_SAM_MAXSEG];
        while (q < e) {
            const char* l1 = (const char*)memchr(q, '\n', (size_t)(e - q));
            if (!l1 || l1 + 1 >= e) { bad[t] = 1; return; }
            const char* l2 = (const char*)memchr(l1 + 1, '\n', (size_t)(e - l1 - 1));
            if (!l2 || *q == '@' || (size_t)(l2 - q) > 0x7FFFFFFFull) { bad[t] = 1; return; }
            ag_sam_parse_line(q, (u32)(l1 - q), a);
            ag_sam_parse_line(l1 + 1, (u32)(l2 - l1 - 1), b);
            if (a.err || b.err) { bad[t] = 1; return; }   // unknown CIGAR character / very long CIGAR: the sequential parser reports or handles it
            PRec r; r.sid = a.sid; r.flags = a.fr | (b.fr << 1); r.p0 = a.tid == AG_NONE ? AG_NONE : ag_sam_pos_at0(a.seg, a.nseg); r.seg_off = (u32)segs[t].size(); r.n1 = r.n2 = 0;
            if (ag_sam_mate_pass(a, kReadThreshold) && ag_sam_mate_pass(b, kReadThreshold)) {
                if (a.tid != 0 || b.tid != 0 || b.sid != a.sid || a.sid >= reads.n_pairs) { bad[t] = 1; return; }
                const u32 rlen = reads.len[a.sid];
                u32 c1 = 0, c2 = 0;
                if (ag_sam_normalize(a.seg, a.nseg, rlen, n1, c1) || ag_sam_normalize(b.seg, b.nseg, rlen, n2, c2) || !c1 || !c2) { bad[t] = 1; return; }
                r.flags |= 4; r.n1 = (uint16_t)c1; r.n2 = (uint16_t)c2;
                segs[t].insert(segs[t].end(), n1, n1 + c1); segs[t].insert(segs[t].end(), n2, n2 + c2);
            }
            recs[t].push_back(r);
            q = l2 + 1;
        }
    };
    { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    for (int t = 0; t < T; t++) if (bad[t]) return false;
    lap("parse threads");
    // ---- order-dependent half, in data-parallel form (the shape a device version takes as well).  Preconditions checked here: read ids
    // non-decreasing.  Then, exactly as the sequential parser behaves on such input:
    //   * batches of 1,000,000 read ids (AG:1885-1894): the first record whose id lies beyond the current batch is consumed and LOST
    //     (AG:1259) and the batch advances by one — a handful of positions, found by binary search on the sorted ids;
    //   * a pair's surviving records form a group (consecutive records that passed AG:1261 with the same id, not separated by a lost
    //     record); a record is dropped when its read-offset-0 position lies within one read length of ANY earlier record of its group
    //     (AG:1650-1655), so every record decides for itself by looking back;
    //   * output order = file order: count per chunk, prefix sums, fill. ----
    std::vector<size_t> base((size_t)T + 1, 0);
    for (int t = 0; t < T; t++) base[t + 1] = base[t] + recs[t].size();
    const size_t N = base[T];
    {   // sorted?
        std::vector<char> unsorted((size_t)T, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            const std::vector<PRec>& r = recs[t];
            for (size_t i = 1; i < r.size(); i++) if (r[i].sid < r[i - 1].sid) { unsorted[t] = 1; return; }
            if (!r.empty()) for (int q = t - 1; q >= 0; q--) if (!recs[q].empty()) { if (r[0].sid < recs[q].back().sid) unsorted[t] = 1; break; } });
        for (auto& x : th) x.join();
        for (int t = 0; t < T; t++) if (unsorted[t]) return false;
    }
    auto first_greater = [&](long v) -> size_t {   // first global index whose id is > v
        for (int t = 0; t < T; t++) {
            const std::vector<PRec>& r = recs[t];
            if (r.empty() || (long)r.back().sid <= v) continue;
            size_t lo = 0, hi = r.size();
            while (lo < hi) { size_t mid = (lo + hi) / 2; if ((long)r[mid].sid > v) hi = mid; else lo = mid + 1; }
            return base[t] + lo;
        }
        return N;
    };
    const long n_pairs = (long)reads.n_pairs;
    std::vector<size_t> lost; size_t stop = N;
    {
        long first = 0, last = std::min<long>(kBatchPairs - 1, n_pairs - 1);
        size_t i = 0;
        for (;;) {
            const size_t j = std::max(i, first_greater(last));
            if (j >= N) break;
            if (last >= n_pairs - 1) { stop = j; break; }   // ids beyond the read set: the reference stops reading here
            lost.push_back(j);
            first = last + 1; last = std::min<long>(first + kBatchPairs - 1, n_pairs - 1);
            i = j + 1;
        }
        (void)first;
    }
    auto is_lost = [&](size_t g) { return std::binary_search(lost.begin(), lost.end(), g); };
    // survives(t, i): passes AG:1261, is neither lost nor beyond `stop`, and no earlier record of its group lies within one read length
    auto survives = [&](int t, size_t i) -> bool {
        const PRec& r = recs[t][i];
        const size_t g = base[t] + i;
        if (!(r.flags & 4) || g >= stop || (!lost.empty() && is_lost(g))) return false;
        u32 rlen = 0; bool have_len = false;
        int qt = t; size_t qi = i;
        for (;;) {   // look back through the group
            if (qi == 0) { do { qt--; } while (qt >= 0 && recs[qt].empty()); if (qt < 0) break; qi = recs[qt].size(); }
            qi--;
            const PRec& o = recs[qt][qi];
            if (!lost.empty() && is_lost(base[qt] + qi)) break;     // a lost record closes the group before it
            if (!(o.flags & 4)) continue;                           // records that failed the filter are invisible to the grouping
            if (o.sid != r.sid) break;
            if (!have_len) { rlen = reads.len[r.sid]; have_len = true; }
            if (absdiff(r.p0, o.p0) < (int)rlen) return false;
        }
        return true;
    };
    std::vector<size_t> n_aln((size_t)T + 1, 0), n_ext((size_t)T + 1, 0);
    std::vector<std::vector<unsigned char>> keep((size_t)T);
    std::vector<char> strand_error((size_t)T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            const std::vector<PRec>& r = recs[t];
            keep[t].assign(r.size(), 0);
            size_t ca = 0, ce = 0;
            for (size_t i = 0; i < r.size(); i++) {
                if (!survives(t, i)) continue;
                const u32 fr1 = r[i].flags & 1, fr2 = (r[i].flags >> 1) & 1;
                if (fr1 == fr2) { strand_error[t] = 1; return; }    // exactly one mate must be reverse (AG:1657-1671)
                keep[t][i] = 1; ca++;
                ce += (r[i].n1 > 1 ? r[i].n1 : 0) + (r[i].n2 > 1 ? r[i].n2 : 0);
            }
            n_aln[t + 1] = ca; n_ext[t + 1] = ce; });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; t++) if (strand_error[t]) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};
    for (int t = 0; t < T; t++) { n_aln[t + 1] += n_aln[t]; n_ext[t + 1] += n_ext[t]; }
    if (n_ext[T] >= 0xFFFFFFF0ull) return false;
    u.aln.resize(n_aln[T]); u.ext.resize(n_ext[T]);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            const std::vector<PRec>& r = recs[t];
            size_t oa = n_aln[t], oe = n_ext[t];
            for (size_t i = 0; i < r.size(); i++) {
                if (!keep[t][i]) continue;
                const PRec& x = r[i];
                const ag_seg* sg = segs[t].data() + x.seg_off;
                ag_aln a; a.pair = x.sid; a.pad = 0;
                a.flags = (x.flags & 3u) | ((u32)x.n1 << 8) | ((u32)x.n2 << 16);
                a.dst1 = sg[0].dst; a.sl1 = sg[0].src | (sg[0].len << 16);
                a.dst2 = sg[x.n1].dst; a.sl2 = sg[x.n1].src | (sg[x.n1].len << 16);
                a.ext_idx = (u32)oe;
                if (x.n1 > 1) { std::copy(sg, sg + x.n1, u.ext.begin() + (long)oe); oe += x.n1; }
                if (x.n2 > 1) { std::copy(sg + x.n1, sg + x.n1 + x.n2, u.ext.begin() + (long)oe); oe += x.n2; }
                u.aln[oa++] = a;
            } });
        for (auto& x : th) x.join();
    }
    lap("merge");
    return true;
}

void ag_parse_sam(const std::string& path, const AgReads& reads, AgUnit& u) {
    FileMap fm(path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    u.aln.clear(); u.ext.clear();
    if (reads.n_pairs == 0) return;
    if (parse_sam_parallel(fm.p, fm.n, reads, u)) return;
    u.aln.clear(); u.ext.clear();
    struct Tmp { u32 pair, fr1, fr2, p0; std::vector<ag_seg> s1, s2; };
    std::vector<Tmp> batch;
    std::vector<ag_seg> segs1, segs2, n1, n2;
    long n_pairs = (long)reads.n_pairs;
    long first = 0, last = std::min<long>(kBatchPairs - 1, n_pairs - 1);
    Lines ln(fm.p, fm.n); const char* s; size_t n;

    auto flush_batch = [&]() {
        bool sorted = true;
        for (size_t i = 1; i < batch.size(); i++) if (batch[i].pair < batch[i - 1].pair) { sorted = false; break; }
        if (!sorted) std::stable_sort(batch.begin(), batch.end(), [](const Tmp& a, const Tmp& b) { return a.pair < b.pair; });
        for (size_t i = 0; i < batch.size();) {
            size_t e = i;
            while (e < batch.size() && batch[e].pair == batch[i].pair) e++;
            u32 rlen = reads.len[batch[i].pair];
            for (size_t pp = i; pp < e; pp++) {
                bool dup = false;  // AG:1650-1655: compared against EVERY earlier alignment of the pair, skipped ones included
                for (size_t q = i; q < pp && !dup; q++) if (absdiff(batch[pp].p0, batch[q].p0) < (int)rlen) dup = true;
                if (dup) continue;
                Tmp& t = batch[pp];
                if (!((t.fr1 == 1 && t.fr2 == 0) || (t.fr2 == 1 && t.fr1 == 0))) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};  // AG:1657-1671
                normalize(t.s1, rlen, n1); normalize(t.s2, rlen, n2);
                if (n1.empty() || n2.empty() || n1.size() > 255 || n2.size() > 255) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};
                ag_aln a; a.pair = t.pair; a.pad = 0;
                a.flags = t.fr1 | (t.fr2 << 1) | ((u32)n1.size() << 8) | ((u32)n2.size() << 16);
                a.dst1 = n1[0].dst; a.sl1 = n1[0].src | (n1[0].len << 16);
                a.dst2 = n2[0].dst; a.sl2 = n2[0].src | (n2[0].len << 16);
                a.ext_idx = (u32)u.ext.size();
                if (n1.size() > 1) u.ext.insert(u.ext.end(), n1.begin(), n1.end());
                if (n2.size() > 1) u.ext.insert(u.ext.end(), n2.begin(), n2.end());
                u.aln.push_back(a);
            }
            i = e;
        }
        batch.clear();
    };

    for (;;) {  // one iteration per read batch (AG:1885-1894)
        bool dropped = false;
        segs1.clear(); segs2.clear();
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '@') continue;
            SamRec a, b;
            parse_sam(s, n, a, segs1);
            if (!ln.next(s, n) || n == 0 || s[0] == 0) throw AgHostError{"BROKEN BOWTIE FILE"};
            parse_sam(s, n, b, segs2);
            if (a.sid < (u32)first) continue;                  // AG:1258 (segment lists are NOT cleared on this path)
            if (a.sid > (u32)last) { dropped = true; break; }  // AG:1259: this record pair is consumed and lost
            if (a.tid != AG_NONE && b.tid != AG_NONE &&
                (double)(a.send - a.sstart - a.sgap) / a.ssize >= kReadThreshold && (double)(a.tend - a.tstart - a.tgap) / (a.tend - a.tstart) >= kReadThreshold &&
                (double)(b.send - b.sstart - b.sgap) / b.ssize >= kReadThreshold && (double)(b.tend - b.tstart - b.tgap) / (b.tend - b.tstart) >= kReadThreshold) {
                if (a.tid != 0 || b.tid != 0 || b.sid != a.sid) throw AgHostError{"BROKEN BOWTIE FILE"};
                Tmp t; t.pair = a.sid; t.fr1 = a.fr; t.fr2 = b.fr; t.p0 = pos_at0(segs1); t.s1 = segs1; t.s2 = segs2;
                batch.push_back(std::move(t));
            }
            segs1.clear(); segs2.clear();
        }
        flush_batch();
        (void)dropped;
        // the reference starts another batch as long as the read file has more pairs (AG:1889-1894), whatever ended this one
        if (last >= n_pairs - 1 || !ln.good) break;
        first = last + 1; last = std::min<long>(first + kBatchPairs - 1, n_pairs - 1);
    }
}

// =============================================================================================================================
// post passes
// =============================================================================================================================
void ag_select_emitted(const std::vector<ag_walk>& walks, std::vector<u32>& sel) {
    sel.clear();
    bool have = false; u32 bso = 0, beo = 0;
    for (size_t i = 0; i < walks.size(); i++) {
        const ag_walk& r = walks[i];
        u32 eoff = r.eoff;
        if (((r.flags >> 1) & 3) != 1) eoff = eoff + (r.tail_soff_len >> 16) - 1;  // AG:2164-2173 (u32 wrap when the string is empty)
        // contain(startIDBak, ..) — the Bak ids are -1 until the first emission, then 0 like every id in a unit (AG:1897-1902)
        if (have && bso <= r.soff && beo >= eoff) continue;
        sel.push_back((u32)i);
        have = true; bso = r.soff; beo = eoff;
    }
}

static inline size_t wrapped_len(size_t n) { return n + (n + 59) / 60; }   // 60 columns, newline after the last base (AG:2179-2184)
static inline char* wrap60_to(char* p, const char* s, size_t n) {
    for (size_t i = 0; i < n; i += 60) { size_t m = std::min<size_t>(60, n - i); memcpy(p, s + i, m); p += m; *p++ = '\n'; }
    return p;
}
// split [0, n) into chunks of roughly equal output bytes for the thread team; cut[i] .. cut[i + 1] is chunk i
static std::vector<size_t> byte_chunks(const std::vector<size_t>& off /* n + 1 prefix offsets */) {
    const size_t n = off.size() - 1, total = off.back();
    int want = std::max(1, std::min<int>(ag_team_size() * 4, (int)(total / (64 << 10)) + 1));
    std::vector<size_t> cut(1, 0);
    for (int c = 1; c < want; c++) {
        size_t target = total / want * c;
        size_t i = (size_t)(std::lower_bound(off.begin(), off.end(), target) - off.begin());
        if (i > cut.back() && i < n) cut.push_back(i);
    }
    cut.push_back(n);
    return cut;
}

void ag_make_contigs_begin(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                           std::vector<AgContig>& contigs, AgMakeState& st) {
    auto tp0 = std::chrono::steady_clock::now();
    const bool probe = getenv("AG_POST_TIMING") != nullptr;
    auto lap = [&](const char* what) { if (probe) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  [make_contigs] . %s %.3f ms\n", what, std::chrono::duration<double>(t - tp0).count() * 1e3); } };
    contigs.clear(); contigs.resize(sel.size());
    // contig records + header strings + where every contig's text goes (nothing here reads the bases).  Headers are formatted by the thread
    // team into fixed slots, then packed.
    std::string& hdr = st.hdr;
    std::vector<size_t>& hoff = st.hoff; std::vector<size_t>& toff = st.toff;
    hoff.assign(sel.size() + 1, 0); toff.assign(sel.size() + 1, 0);
    const size_t SLOT = 160;
    static thread_local std::vector<char> slots;   // (one materialisation at a time per calling thread)
    std::vector<unsigned char> hlen(sel.size());
    slots.resize(sel.size() * SLOT);
    char* const slot_base = slots.data();   // (the team's threads must not name the thread_local themselves: each would see its own, empty one)
    lap("sized");
    const int nchunk = (int)std::min<size_t>((sel.size() + 63) / 64, (size_t)std::max(1, ag_team_size() * 2));
    const size_t per = nchunk ? (sel.size() + (size_t)nchunk - 1) / (size_t)nchunk : 0;
    ag_parallel_chunks(nchunk, [&](int ch) {
      for (size_t i = (size_t)ch * per; i < std::min(sel.size(), ((size_t)ch + 1) * per); i++) {
        const ag_walk& r = walks[sel[i]];
        AgContig& c = contigs[i];
        c.extended = (int)(r.flags & 1);
        c.sid = 0; c.soff = r.soff; c.eid = 0; c.eoff = r.eoff;
        c.sid0 = r.soff0 == AG_NONE ? AG_NONE : 0; c.soff0 = r.soff0;
        c.p = bases + offs[i]; c.n = (size_t)(offs[i + 1] - offs[i]);
        u32 mode = (r.flags >> 1) & 3;
        if (mode == 1) { c.eid0 = AG_NONE; c.eoff0 = AG_NONE; }  // walk ended on a contiMer (AG:2158-2162)
        else {
            c.eid0 = r.eoff0 == AG_NONE ? AG_NONE : 0; c.eoff0 = r.eoff0;
            const u32 slen = r.tail_soff_len >> 16;
            c.eoff = c.eoff + slen - 1; c.eoff0 = c.eoff0 + slen - 1;   // size_t arithmetic truncated to u32 (AG:2170-2171)
        }
        {   // ">i, extended, sid, soff, eid, eoff, sid0, soff0, eid0, eoff0 \n"  (AG:2178)
            char* hb = slot_base + i * SLOT; char* w = hb;
            auto num = [&](unsigned long v) { char t[24]; int k = 24; do { t[--k] = (char)('0' + v % 10); v /= 10; } while (v); memcpy(w, t + k, (size_t)(24 - k)); w += 24 - k; };
            auto sep = [&]() { *w++ = ','; *w++ = ' '; };
            *w++ = '>'; num(i); sep();
            if (c.extended < 0) { *w++ = '-'; num((unsigned long)(-(long)c.extended)); } else num((unsigned long)c.extended);
            sep(); num(c.sid); sep(); num(c.soff); sep(); num(c.eid); sep(); num(c.eoff); sep(); num(c.sid0); sep(); num(c.soff0); sep(); num(c.eid0); sep(); num(c.eoff0);
            *w++ = ' '; *w++ = '\n';
            hlen[i] = (unsigned char)(w - hb);
        }
      }
    });
    lap("records");
    for (size_t i = 0; i < sel.size(); i++) { hoff[i + 1] = hoff[i] + hlen[i]; toff[i + 1] = toff[i] + hlen[i] + wrapped_len((size_t)(offs[i + 1] - offs[i])); }
    hdr.resize(hoff.back());
    ag_parallel_chunks(nchunk, [&](int ch) {
        for (size_t i = (size_t)ch * per; i < std::min(sel.size(), ((size_t)ch + 1) * per); i++) memcpy(&hdr[hoff[i]], slot_base + i * SLOT, hlen[i]);
    });
    if (getenv("AG_POST_TIMING")) fprintf(stderr, "  [make_contigs] records + headers %.2f ms\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count() * 1e3);
}

void ag_make_contigs_finish(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                            const AgReads& reads, const std::vector<AgContig>& contigs, const AgMakeState& st, AgText& pre_text) {
    auto tp1 = std::chrono::steady_clock::now();
    if (!reads.exc.empty())   // the device wrote 'N' for every masked base of a tail; put the original characters back (AG:2167 copies s verbatim)
        for (size_t i = 0; i < sel.size(); i++) {
            const ag_walk& r = walks[sel[i]];
            if (((r.flags >> 1) & 3) == 1) continue;
            const u32 slen = r.tail_soff_len >> 16, soff = r.tail_soff_len & 0xFFFFu, read = r.tail_sread >> 1, rc = r.tail_sread & 1;
            if (slen <= 1) continue;
            const u32 rlen = reads.len[read >> 1];
            char* t = bases + offs[i] + r.len;
            for (u32 j = 1; j < slen; j++) { const char c = reads.at(read, rc, rlen, soff + j); if (c) t[j - 1] = c; }
        }
    // thread team: header + 60-column body of every contig at its offset
    const std::vector<size_t>& hoff = st.hoff; const std::vector<size_t>& toff = st.toff;
    pre_text.set_size(toff.back());
    const std::vector<size_t> cut = byte_chunks(toff);
    char* const T = pre_text.data();
    ag_parallel_chunks((int)cut.size() - 1, [&](int ch) {
        for (size_t i = cut[ch]; i < cut[ch + 1]; i++) {
            char* p = T + toff[i];
            memcpy(p, st.hdr.data() + hoff[i], hoff[i + 1] - hoff[i]); p += hoff[i + 1] - hoff[i];
            wrap60_to(p, contigs[i].p, contigs[i].n);
        }
    });
    if (getenv("AG_POST_TIMING")) fprintf(stderr, "  [make_contigs] fill %.2f ms (%d chunks, team %d)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp1).count() * 1e3, (int)cut.size() - 1, ag_team_size());
}

void ag_make_contigs(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char* bases, const std::vector<u64>& offs,
                     const AgReads& reads, std::vector<AgContig>& contigs, AgText& pre_text) {
    AgMakeState st;
    ag_make_contigs_begin(walks, sel, bases, offs, contigs, st);
    ag_make_contigs_finish(walks, sel, bases, offs, reads, contigs, st, pre_text);
}

static inline int contain(const AgContig& a, const AgContig& b) { return a.sid == b.sid && a.eid == b.eid && a.soff <= b.soff && a.eoff >= b.eoff; }

void ag_dedup_join(std::vector<AgContig>& cs) {
    struct Lap { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(); ~Lap() { if (getenv("AG_POST_TIMING")) fprintf(stderr, "  [dedup_join] %.2f ms\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3); } } lap;
    int n = (int)cs.size();
    for (int a = 0; a < n; a++) {  // forward containment (AG:2303-2320)
        if (cs[a].extended != 1) continue;
        for (int b = a + 1; b < n; b++) {
            if (contain(cs[a], cs[b])) cs[b].extended = 2;
            else if (cs[a].eid != cs[b].sid || cs[a].eoff < cs[b].soff) break;
        }
    }
    for (int a = n - 1; a >= 0; a--) {  // backward containment (AG:2322-2339)
        if (cs[a].extended != 1) continue;
        for (int b = a - 1; b >= 0; b--) {
            if (contain(cs[a], cs[b])) cs[b].extended = 2;
            else if (cs[b].eid != cs[a].sid || cs[b].eoff < cs[a].soff) break;
        }
    }
    for (int a = 0; a < n; a++) {  // join with the single later contig that starts inside (AG:2342-2378)
        while (cs[a].extended == 1) {
            int hits = 0, lastb = -1;
            for (int b = a + 1; b < n; b++) {
                if (cs[b].extended == 2) continue;
                if (cs[a].eoff >= cs[b].soff) { hits++; lastb = b; } else break;
            }
            if (hits != 1) break;
            AgContig& t = cs[lastb];
            t.extended = 2;
            u32 from = cs[a].eoff - t.soff + 1;
            if (from < t.size()) t.pieces_from(from, cs[a].more);   // the reference appends t's suffix (AG:2368-2370); here only views
            cs[a].eid = t.eid; cs[a].eoff = t.eoff; cs[a].eid0 = t.eid0; cs[a].eoff0 = t.eoff0;
        }
    }
}

static inline int overlap(u32 x1, u32 y1, u32 x2, u32 y2) {  // AG:2388-2394
    return (x1 <= x2 && x2 <= y1 && y1 <= y2 && (int)y1 - (int)x2 > 0) || (x2 <= x1 && x1 <= y2 && y2 <= y1 && (int)y2 - (int)x1 > 0) ||
           (x1 <= x2 && x2 <= y2 && y2 <= y1 && (int)y2 - (int)x2 > 0) || (x2 <= x1 && x1 <= y1 && y1 <= y2 && (int)y1 - (int)x1 > 0);
}

void ag_scaffold(std::vector<AgContig>& cs, const std::string& ref, const std::vector<unsigned char>& occ, AgText& text) {
    auto ts0 = std::chrono::steady_clock::now();
    // a scaffold = a list of pieces (contig sequences and reference gap fills); nothing is copied until the text is written
    typedef AgPiece Piece;
    std::vector<Piece> pieces; std::vector<size_t> first(1, 0);   // pieces of scaffold i: [first[i], first[i + 1])
    auto occupied = [&](u32 p) { return (size_t)(p >> 3) < occ.size() && ((occ[p >> 3] >> (p & 7)) & 1); };
    // the reference scans EVERY later contig for a chaining partner (quadratic); only contigs with extended == 1 can ever satisfy the
    // test and `extended` does not change here, so scanning that sub-list in the same order is equivalent
    std::vector<u32> ext1;
    for (u32 i = 0; i < cs.size(); i++) if (cs[i].extended == 1) ext1.push_back(i);
    for (u32 cp = 0; cp < cs.size(); cp++) {
        if (!(cs[cp].sid != AG_NONE && cs[cp].extended == 1)) continue;
        cs[cp].pieces_from(0, pieces);
        cs[cp].sid = AG_NONE;
        int cont = 1;
        while (cs[cp].sid0 == cs[cp].eid0 && cont) {
            cont = 0;
            for (size_t e = (size_t)(std::upper_bound(ext1.begin(), ext1.end(), cp) - ext1.begin()); e < ext1.size(); e++) {
                const u32 c0 = ext1[e];
                const AgContig& a = cs[cp]; AgContig& b = cs[c0];
                // contigs are in start-position order and every case of overlap() needs b.soff <= a.eoff0: nothing further can match
                if (a.eoff0 != AG_NONE && b.soff != AG_NONE && b.soff > a.eoff0 && b.sid != AG_NONE) break;
                if (!(a.eid0 == b.sid && b.sid == b.eid && overlap(a.soff0, a.eoff0, b.soff, b.eoff) && b.extended == 1)) continue;
                if (b.soff > a.eoff) {
                    u32 gap = b.soff - a.eoff - 1; int covered = 0;
                    for (u32 i = 0; i < gap; i++) if (occupied(a.eoff + i + 1)) covered++;
                    if ((gap != 0 && (double)covered / gap >= 0.5) || gap == 0) { if (gap) pieces.push_back(Piece{ref.data() + a.eoff + 1, gap}); }   // reference bases fill the gap (AG:2428-2436)
                    else continue;
                }
                b.pieces_from(0, pieces);
                b.sid = AG_NONE;
                cp = c0; cont = 1;
                break;
            }
        }
        first.push_back(pieces.size());
    }
    const size_t ns = first.size() - 1;
    auto ts1 = std::chrono::steady_clock::now();
    std::string hdr; std::vector<size_t> hoff(ns + 1, 0), toff(ns + 1, 0);
    { Out out(&hdr); for (size_t i = 0; i < ns; i++) {
        out.ch('>'); out.num(i); out.ch('\n'); hoff[i + 1] = hdr.size();
        size_t n = 0; for (size_t k = first[i]; k < first[i + 1]; k++) n += pieces[k].n;
        toff[i + 1] = toff[i] + (hoff[i + 1] - hoff[i]) + wrapped_len(n);
    } }
    text.set_size(toff.back());
    const std::vector<size_t> cut = byte_chunks(toff);
    char* const T = text.data();
    ag_parallel_chunks((int)cut.size() - 1, [&](int ch) {
        for (size_t i = cut[ch]; i < cut[ch + 1]; i++) {
            char* p = T + toff[i];
            memcpy(p, hdr.data() + hoff[i], hoff[i + 1] - hoff[i]); p += hoff[i + 1] - hoff[i];
            if (first[i + 1] - first[i] == 1) { wrap60_to(p, pieces[first[i]].p, pieces[first[i]].n); continue; }
            size_t col = 0;   // 60-column wrapping across the pieces
            for (size_t k = first[i]; k < first[i + 1]; k++) {
                const char* q = pieces[k].p; size_t n = pieces[k].n;
                while (n) { size_t m = std::min(n, 60 - col); memcpy(p, q, m); p += m; q += m; n -= m; col += m; if (col == 60) { *p++ = '\n'; col = 0; } }
            }
            if (col) *p++ = '\n';
        }
    });
    if (getenv("AG_POST_TIMING")) fprintf(stderr, "  [scaffold] chaining %.2f ms, text %.2f ms (%zu scaffolds)\n", std::chrono::duration<double>(ts1 - ts0).count() * 1e3,
                                          std::chrono::duration<double>(std::chrono::steady_clock::now() - ts1).count() * 1e3, ns);
}

// =============================================================================================================================
// input normalisation (--resume re-runs these, AG:4757-4758)
// =============================================================================================================================
// formalizeInput (contigs), AG:3228-3320: chunks of at most LARGE_CHUNK bases named ">chunk.contig"; contigs of <= 200 bases go to the chaff
// file (only tmp/_contigs.fa has one: for tmp/_<id>_contigs.fa the reference writes them to a closed stream, i.e. drops them)
void ag_formalize_contigs_to(const std::string& in_path, const std::string& out_path, const std::string& chaff_path, std::vector<std::string>& contig_ids) {
    FileMap fm(in_path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    std::vector<std::string> seqs, ids;
    {
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') { seqs.emplace_back(); ids.emplace_back(s + 1, n - 1); }
            else if (!seqs.empty()) seqs.back().append(s, n);
        }
    }
    Out out(out_path);
    std::unique_ptr<Out> chaff;
    if (!chaff_path.empty()) chaff.reset(new Out(chaff_path));
    contig_ids.clear();
    unsigned long chunk = 0, real = 0;
    for (size_t c = 0; c < seqs.size(); c++) {
        const std::string& q = seqs[c];
        if (q.size() > 200) {
            auto header = [&]() { out.ch('>'); out.num(chunk++); out.ch('.'); out.num(real); out.ch('\n'); };
            header();
            if ((long)q.size() < kLargeChunk) out.wrap60(q);
            else {
                long total = 0;
                for (long i = 0; i < (long)q.size(); i++) {
                    out.ch(q[i]);
                    if ((i + 1) % kLargeChunk == 0 && i < (long)q.size() - 1 - 60) { total += kLargeChunk; out.ch('\n'); header(); continue; }
                    if ((i + 1 - total) % 60 == 0 || i == (long)q.size() - 1) out.ch('\n');
                }
            }
            real++;
            contig_ids.push_back(ids[c]);
        } else if (chaff) { chaff->ch('>'); chaff->put(ids[c]); chaff->ch('\n'); chaff->wrap60(q); }
    }
}
void ag_formalize_contigs(const std::string& in_path, const std::string& tmp, std::vector<std::string>& contig_ids) {
    ag_formalize_contigs_to(in_path, tmp + "/_contigs.fa", tmp + "/_chaff.fa", contig_ids);
}

int ag_formalize_genome(const std::string& in_path, const std::string& tmp, int part, std::vector<std::string>& genome_ids) {
    FileMap fm(in_path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    std::vector<std::string> chr;
    {
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') { chr.emplace_back(); genome_ids.emplace_back(s + 1, n - 1); }
            else if (!chr.empty()) chr.back().append(s, n);
        }
    }
    int unit = 0;
    Out all(tmp + "/_genome.fa");
    for (size_t g = 0; g < chr.size(); g++) {
        const std::string& q = chr[g];
        Out* out = new Out(tmp + "/_genome." + std::to_string(unit) + ".fa");
        out->put(">0\n", 3); all.ch('>'); all.num((unsigned long)unit); all.ch('\n');
        int k = 1; long slice = (long)q.size() / part;
        for (long i = 0; i < (long)q.size(); i++) {
            out->ch(q[i]); all.ch(q[i]);
            bool cut = slice > 0 && ((i + 1) % slice == 0 && k < part);
            if ((i + 1) % 60 == 0 || i == (long)q.size() - 1 || cut) { out->ch('\n'); all.ch('\n'); }
            if (i != (long)q.size() - 1 && cut) {
                delete out; unit++; k++;
                out = new Out(tmp + "/_genome." + std::to_string(unit) + ".fa");
                out->put(">0\n", 3); all.ch('>'); all.num((unsigned long)unit); all.ch('\n');
            }
        }
        delete out; unit++;
    }
    return unit;
}

// One open / fallocate / write / close per output file.  Reserving the extent first (fallocate; ignored where the file system has no
// support) lets ext4 allocate the blocks once instead of reserving them page by page on the buffered-write path: about twice as fast for
// the 5 MB per-unit FASTA files here.  AG_WRITE_PLAIN=1 skips the reservation (measurements).
static void write_file_bytes(const std::string& path, const char* data, size_t n) {
    static const bool plain = getenv("AG_WRITE_PLAIN") != nullptr;
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) throw AgHostError{"CANNOT OPEN FILE!"};
    if (n >= ((size_t)1 << 20) && !plain) (void)!fallocate(fd, 0, 0, (off_t)n);
    size_t a = 0;
    while (a < n) {
        const ssize_t got = write(fd, data + a, n - a);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) break;          // (disk full, ...: like the reference's unchecked ofstream, the file stays short)
        a += (size_t)got;
    }
    close(fd);
}
void ag_write_file(const std::string& path, const AgText& text) { write_file_bytes(path, text.data(), text.size()); }
void ag_write_file(const std::string& path, const std::string& text) { write_file_bytes(path, text.data(), text.size()); }

// =============================================================================================================================
// CLI-side phases outside the hot path (kept byte-compatible with the reference so that the binary is a drop-in)
// =============================================================================================================================
int ag_max_read_length(const std::string& path) {  // AG:3197-3226
    FileMap fm(path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    int mx = 0, len = 0;
    Lines ln(fm.p, fm.n); const char* s; size_t n;
    while (ln.next(s, n)) {
        if (n == 0 || s[0] == 0) break;
        if (s[0] == '>') { if (len > mx) mx = len; len = 0; continue; }
        len += (int)n;
    }
    if (len > mx) mx = len;
    return mx;
}

// AG:3420-3518: rename pairs to integers, truncate both mates to the shorter one, write the interleaved and the two split files
long ag_formalize_reads(const std::string& in1, const std::string& in2, const std::string& tmp) {
    FileMap f1(in1), f2(in2);
    if (!f1.ok || !f2.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    Out out(tmp + "/_reads.fa"), out1(tmp + "/_reads_1.fa"), out2(tmp + "/_reads_2.fa");
    Lines l1(f1.p, f1.n), l2(f2.p, f2.n);
    std::string r1, r2;
    unsigned long id = 0;
    auto emit = [&]() {
        if (r1.empty() || r2.empty()) return;
        size_t sz = std::min(r1.size(), r2.size());
        for (Out* o : {&out, &out1}) { o->ch('>'); o->num(id); o->ch('\n'); o->put(r1.data(), sz); o->ch('\n'); }
        for (Out* o : {&out, &out2}) { o->ch('>'); o->num(id); o->ch('\n'); o->put(r2.data(), sz); o->ch('\n'); }
        id++;
    };
    for (;;) {
        const char *s1 = nullptr, *s2 = nullptr; size_t n1 = 0, n2 = 0;
        bool g1 = l1.next(s1, n1), g2 = l2.next(s2, n2);
        if (!g1 || !g2) break;  // while(in1.good() && in2.good())
        bool e1 = n1 == 0 || s1[0] == 0, e2 = n2 == 0 || s2[0] == 0;
        if (e1 && e2) break;
        if (e1 != e2) throw AgHostError{"INCONSISTENT PE FILES!"};
        if (s1[0] == '>' && s2[0] == '>') { emit(); r1.clear(); r2.clear(); }
        else if (s1[0] != '>' && s2[0] != '>') { r1.append(s1, n1); r2.append(s2, n2); }
        else throw AgHostError{"INCONSISTENT PE FILES!"};
    }
    emit();
    return (long)id;
}

// AG:3545-3579 (+ parseBT AG:3520-3543): split the whole-genome SAM by the integer RNAME
void ag_distribute_alignments(const std::string& tmp, int units) {
    FileMap fm(tmp + "/_reads_genome.bowtie");
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    std::vector<Out*> outs;
    for (int u = 0; u < units; u++) outs.push_back(new Out(tmp + "/_reads_genome." + std::to_string(u) + ".bowtie"));
    Lines ln(fm.p, fm.n); const char* s; size_t n;
    while (ln.next(s, n)) {
        if (n && s[0] == '@') continue;
        if (n == 0 || s[0] == 0) break;
        const char* end = s + n; const char* p = s; int tab = 0;
        while (p < end && tab < 2) { if (*p == '\t') tab++; p++; }
        const char* q = p;
        while (q < end && *q != '\t') q++;
        int target = memchr(p, '*', (size_t)(q - p)) ? -1 : ag_atoi(p, std::min<size_t>((size_t)(q - p), 9));
        if (target >= 0 && target < units) { outs[target]->put(s, n); outs[target]->ch('\n'); }
    }
    for (Out* o : outs) delete o;
}

// AG:3751-3819
double ag_check_ratio(const std::string& tmp, int units) {
    long n_reads = 0;
    {
        FileMap fm(tmp + "/_reads_1.fa");
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        for (size_t i = 0; i < fm.n; i++) if (fm.p[i] == '>' && (i == 0 || fm.p[i - 1] == '\n')) n_reads++;
    }
    std::vector<char> hit((size_t)n_reads, 0);
    std::vector<ag_seg> s1, s2;
    for (int u = 0; u < units; u++) {
        FileMap fm(tmp + "/_reads_genome." + std::to_string(u) + ".bowtie");
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '@') continue;
            SamRec a, b;
            parse_sam(s, n, a, s1);
            if (!ln.next(s, n) || n == 0 || s[0] == 0) throw AgHostError{"BROKEN BOWTIE FILE"};
            parse_sam(s, n, b, s2);
            if (a.tid != AG_NONE && b.tid != AG_NONE &&
                (double)(a.send - a.sstart - a.sgap) / a.ssize >= kReadThreshold && (double)(a.tend - a.tstart - a.tgap) / (a.tend - a.tstart) >= kReadThreshold &&
                (double)(b.send - b.sstart - b.sgap) / b.ssize >= kReadThreshold && (double)(b.tend - b.tstart - b.tgap) / (b.tend - b.tstart) >= kReadThreshold)
                if ((long)a.sid < n_reads) hit[a.sid] = 1;
        }
    }
    long aligned = 0;
    for (char c : hit) aligned += c;
    return aligned == 0 ? 0.0 : (double)aligned / (double)n_reads;
}

// AG:2864-3195.  `blat` runs the aligner for unit i (database tmp/_extended_contigs.i.fa, query tmp/_short_initial_contigs.i.fa,
// output tmp/_short_initial_contigs_extended_contigs.i.psl) and returns false when both pblat and blat failed.
void ag_refinement(const std::string& tmp, int units, const std::vector<std::string>& genome_ids, const std::vector<std::string>& contig_ids,
                   int unique_extension, const std::string& ext_path, const std::string& rmn_path, bool (*blat)(int unit, void* user), void* user,
                   bool write_test_files) {
    const size_t kSmallChunk = 20000;  // SMALL_CHUNK, AG:41
    auto load_fasta = [](const std::string& path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
        FileMap fm(path);
        if (!fm.ok) return false;
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') { names.emplace_back(s + 1, n - 1); seqs.emplace_back(); }
            else if (!seqs.empty()) seqs.back().append(s, n);
        }
        return true;
    };
    for (int i = 0; i < units; i++) {  // truncate the initial contigs to 20 kbp for the aligner (AG:2891-2953)
        std::vector<std::string> names, seqs;
        if (!load_fasta(tmp + "/_initial_contigs." + std::to_string(i) + ".fa", names, seqs)) throw AgHostError{"CANNOT OPEN FILE!"};
        Out o(tmp + "/_short_initial_contigs." + std::to_string(i) + ".fa");
        for (size_t j = 0; j < seqs.size(); j++) {
            int num = ag_atoi(names[j].data(), names[j].size());
            o.ch('>'); o.inum(num);
            if (seqs[j].size() > kSmallChunk) { o.ch('.'); o.num(seqs[j].size()); o.ch('\n'); o.wrap60(seqs[j].substr(0, kSmallChunk)); }
            else { o.ch('\n'); o.wrap60(seqs[j]); }
        }
    }
    for (int i = 0; i < units; i++) if (!blat(i, user)) throw AgHostError{"BLAT CALL FAILED!"};
    // original contigs regrouped from the chunks (AG:2985-3014)
    std::vector<std::string> init;
    {
        FileMap fm(tmp + "/_contigs.fa");
        if (!fm.ok) { printf("CANNOT OPEN FILE!\n"); return; }
        Lines ln(fm.p, fm.n); const char* s; size_t n; int prev = -1;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') {
                const char* dot = (const char*)memchr(s, '.', n);
                int id = dot ? ag_atoi(dot + 1, (size_t)(s + n - dot - 1)) : 0;
                if (id != prev) { init.emplace_back(); prev = id; }
            } else if (!init.empty()) init.back().append(s, n);
        }
    }
    std::vector<int> init_tags(init.size(), 0);
    std::vector<std::vector<int>> extd_init_map;  // NOT reset between units in the reference (AG:2877, AG:3035)
    Out e(ext_path), r(rmn_path);
    Out* test_in = write_test_files ? new Out(std::string("in.fa")) : nullptr;
    Out* test_ex = write_test_files ? new Out(std::string("ex.fa")) : nullptr;
    unsigned long seq_id = 0;
    for (int i = 0; i < units; i++) {
        std::vector<std::string> names, extd;
        if (!load_fasta(tmp + "/_extended_contigs." + std::to_string(i) + ".fa", names, extd)) { printf("CANNOT OPEN FILE!\n"); return; }
        std::vector<int> extd_tags(extd.size(), 0);
        for (size_t j = 0; j < extd.size(); j++) extd_init_map.emplace_back();
        int target_bak = -1;
        FileMap fm(tmp + "/_short_initial_contigs_extended_contigs." + std::to_string(i) + ".psl");
        if (!fm.ok) { printf("CANNOT OPEN FILE!\n"); return; }
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        PslRec p;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            parse_psl(s, n, p);
            // parseBLAT's return value and targetSize (AG:501-521)
            const char* f = s; int tab = 0; const char* fe = s + n; const char* q9 = nullptr; size_t l9 = 0; u32 tsize = 0;
            for (const char* c = s; c <= fe; c++) {
                if (c == fe || *c == '\t') { if (tab == 9) { q9 = f; l9 = (size_t)(c - f); } if (tab == 14) tsize = (u32)ag_atoi(f, (size_t)(c - f)); tab++; f = c + 1; }
            }
            u32 real_size = p.ssize;
            if (q9) { const char* dot = (const char*)memchr(q9, '.', l9); if (dot) real_size = (u32)ag_atoi(dot + 1, (size_t)(q9 + l9 - dot - 1)); }
            if (!((double)(p.send - p.sstart - p.sgap) / p.ssize >= 0.8 && (double)(p.tend - p.tstart - p.tgap) / (double)(p.tend - p.tstart) >= 0.8 &&
                  tsize > real_size + 100 && real_size > tsize / 100)) continue;
            if (p.tid >= extd_tags.size() || p.sid >= init_tags.size()) continue;  // the reference would index out of bounds
            if (unique_extension == 1) {
                if (init_tags[p.sid] > 0 && target_bak != -1) {
                    if (extd_tags[target_bak] < extd_tags[p.tid]) {
                        extd_tags[target_bak] = 0;
                        if (!extd_init_map[target_bak].empty()) extd_init_map[target_bak].pop_back();
                        extd_tags[p.tid] = (int)tsize; init_tags[p.sid] = 1; extd_init_map[p.tid].push_back((int)p.sid);
                    }
                } else { extd_tags[p.tid] = (int)tsize; init_tags[p.sid] = 1; extd_init_map[p.tid].push_back((int)p.sid); }
                target_bak = (int)p.tid;
            } else { extd_tags[p.tid] = 1; init_tags[p.sid] = 1; extd_init_map[p.tid].push_back((int)p.sid); }
        }
        for (size_t j = 0; j < extd_tags.size(); j++) {
            if (extd_tags[j] <= 0) continue;
            // genomeIds[i] with i = unit index (AG:3102); with --part > 1 the reference reads past the vector —
            // we stay in bounds and name the last chromosome instead of crashing
            const std::string& gid = genome_ids.empty() ? std::string() : genome_ids[std::min<size_t>((size_t)i, genome_ids.size() - 1)];
            e.put(">AlignGraph", 11); e.num(seq_id); e.put(" @ ", 3); e.put(gid); e.put(" : ", 3);
            for (int c : extd_init_map[j]) { if ((size_t)c < contig_ids.size()) e.put(contig_ids[(size_t)c]); e.put(" ; ", 3); }
            e.ch('\n');
            if (test_ex) { test_ex->ch('>'); test_ex->inum(i); test_ex->put(": ", 2); test_ex->num(seq_id); test_ex->ch('\n'); test_ex->wrap60(extd[j]); }
            seq_id++;
            e.wrap60(extd[j]);
        }
    }
    for (size_t i = 0; i < init_tags.size(); i++) {
        if (init_tags[i] != 0) continue;
        r.ch('>'); if (i < contig_ids.size()) r.put(contig_ids[i]); r.ch('\n'); r.wrap60(init[i]);
    }
    {
        FileMap fm(tmp + "/_chaff.fa");
        if (!fm.ok) { printf("CANNOT OPEN FILE!\n"); }
        else { Lines ln(fm.p, fm.n); const char* s; size_t n; while (ln.next(s, n)) { if (n == 0 || s[0] == 0) break; r.put(s, n); r.ch('\n'); } }
    }
    if (test_in) for (size_t i = 0; i < init_tags.size(); i++) if (init_tags[i] == 1) { test_in->ch('>'); test_in->num(i); test_in->ch('\n'); test_in->wrap60(init[i]); }
    delete test_in; delete test_ex;
}



// =============================================================================================================================
// built-in containment search for refinement() when no BLAT is installed (see ag_host.h)
// =============================================================================================================================
namespace {
void load_seqset(const std::string& path, AgSeqSet& out) {
    FileMap fm(path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    out.names.clear(); out.off.clear(); out.blob.clear();
    Lines ln(fm.p, fm.n); const char* s; size_t n;
    while (ln.next(s, n)) {
        if (n == 0) { if (!ln.good) break; continue; }
        if (s[0] == '>') { out.names.emplace_back(s + 1, n - 1); out.off.push_back(out.blob.size()); }
        else if (!out.names.empty()) out.blob.append(s, n);
    }
    out.off.push_back(out.blob.size());
}
inline char comp_c(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
inline u64 hash24(const char* p) { u64 h = 1469598103934665603ull; for (int i = 0; i < 24; i++) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; } return h; }
}  // namespace

void ag_verify_placements_host(const AgSeqSet& db, const AgSeqSet& qs, const std::vector<AgPlacement>& cand, std::vector<u32>& match, void*) {
    match.assign(cand.size(), 0);
    for (size_t c = 0; c < cand.size(); c++) {
        const AgPlacement& p = cand[c];
        const char* q = qs.blob.data() + qs.off[p.q]; const size_t len = qs.len(p.q);
        const char* t = db.blob.data() + db.off[p.t] + p.start;
        u32 m = 0;
        for (size_t i = 0; i < len; i++) m += (p.strand ? comp_c(q[len - 1 - i]) : q[i]) == t[i];
        match[c] = m;
    }
}

void ag_contain_search(const std::string& db_fa, const std::string& query_fa, const std::string& out_psl, AgVerifyFn verify, void* user) {
    const size_t K = 24;
    AgSeqSet db, qs;
    load_seqset(db_fa, db); load_seqset(query_fa, qs);
    // index of the database 24-mers: (hash, sequence, offset), sorted
    struct Ent { u64 h; u32 t; u32 i; };
    std::vector<Ent> idx;
    for (size_t t = 0; t < db.n(); t++) { const char* s = db.blob.data() + db.off[t]; const size_t L = db.len(t); for (size_t i = 0; i + K <= L; i++) idx.push_back(Ent{hash24(s + i), (u32)t, (u32)i}); }
    std::sort(idx.begin(), idx.end(), [](const Ent& a, const Ent& b) { return a.h != b.h ? a.h < b.h : a.t != b.t ? a.t < b.t : a.i < b.i; });
    auto hits = [&](const char* kmer, std::vector<std::pair<u32, u32>>& out) {   // occurrences of the 24-mer, in (sequence, offset) order
        out.clear();
        const u64 h = hash24(kmer);
        auto it = std::lower_bound(idx.begin(), idx.end(), h, [](const Ent& a, u64 v) { return a.h < v; });
        for (; it != idx.end() && it->h == h; ++it) if (memcmp(db.blob.data() + db.off[it->t] + it->i, kmer, K) == 0) out.push_back({it->t, it->i});
    };
    // oriented queries and their full-length candidates
    std::vector<AgPlacement> cand; std::vector<size_t> first;   // candidates of (query, strand) = [first[2q + strand], first[2q + strand + 1])
    std::vector<std::string> oriented(2 * qs.n());
    std::vector<std::pair<u32, u32>> hv;
    for (size_t q = 0; q < qs.n(); q++)
        for (u32 strand = 0; strand < 2; strand++) {
            first.push_back(cand.size());
            std::string& s = oriented[2 * q + strand];
            s.assign(qs.blob.data() + qs.off[q], qs.len(q));
            if (strand) { std::reverse(s.begin(), s.end()); for (char& c : s) c = comp_c(c); }
            if (s.size() < K) continue;
            std::vector<std::pair<u32, long>> seen;
            for (size_t off = 0; off + K <= s.size(); off += std::max<size_t>(K, s.size() / 16)) {
                hits(s.data() + off, hv);
                for (auto& h : hv) {
                    const long start = (long)h.second - (long)off;
                    if (start < 0 || start + (long)s.size() > (long)db.len(h.first)) continue;
                    if (std::find(seen.begin(), seen.end(), std::make_pair(h.first, start)) != seen.end()) continue;
                    seen.push_back({h.first, start});
                    cand.push_back(AgPlacement{(u32)q, strand, h.first, start});
                }
            }
        }
    first.push_back(cand.size());
    std::vector<u32> match;
    verify(db, qs, cand, match, user);
    Out out(out_psl);
    char line[512];
    for (size_t q = 0; q < qs.n(); q++)
        for (u32 strand = 0; strand < 2; strand++) {
            const std::string& s = oriented[2 * q + strand];
            if (s.size() < K) continue;
            bool placed = false;
            for (size_t c = first[2 * q + strand]; c < first[2 * q + strand + 1]; c++) {
                const long m = match[c];
                if (m * 10 < (long)s.size() * 9) continue;
                placed = true;
                const AgPlacement& p = cand[c];
                int n = snprintf(line, sizeof line, "%ld\t%ld\t0\t0\t0\t0\t0\t0\t%c\t%s\t%zu\t0\t%zu\t%s\t%zu\t%ld\t%ld\t1\t%zu,\t0,\t%ld,\n", m, (long)s.size() - m, strand ? '-' : '+',
                                 qs.names[q].c_str(), s.size(), s.size(), db.names[p.t].c_str(), db.len(p.t), p.start, p.start + (long)s.size(), s.size(), p.start);
                out.put(line, (size_t)n);
            }
            if (placed) continue;
            // local ungapped alignments: seeds every K bases, every new diagonal extended both ways with an X-drop (match +1, mismatch -3, drop 30)
            std::vector<std::pair<std::pair<u32, long>, long>> done;
            for (size_t off = 0; off + K <= s.size(); off += K) {
                hits(s.data() + off, hv);
                for (auto& h : hv) {
                    const u32 t = h.first; const long diag = (long)h.second - (long)off; const char* ts = db.blob.data() + db.off[t]; const long tl = (long)db.len(t);
                    long qa = (long)off, qe = (long)off + (long)K;
                    { long score = 0, best = 0, i = qe, be = qe;
                      while (i < (long)s.size() && i + diag < tl) { score += s[(size_t)i] == ts[i + diag] ? 1 : -3; i++; if (score > best) { best = score; be = i; } if (score < best - 30) break; }
                      qe = be; }
                    { long score = 0, best = 0, i = qa - 1, bs = qa;
                      while (i >= 0 && i + diag >= 0) { score += s[(size_t)i] == ts[i + diag] ? 1 : -3; if (score > best) { best = score; bs = i; } if (score < best - 30) break; i--; }
                      qa = bs; }
                    if (qe - qa < 100) continue;
                    const auto key = std::make_pair(std::make_pair(t, diag), qa);
                    if (std::find(done.begin(), done.end(), key) != done.end()) continue;
                    done.push_back(key);
                    long m = 0;
                    for (long i = qa; i < qe; i++) m += s[(size_t)i] == ts[i + diag];
                    int n = snprintf(line, sizeof line, "%ld\t%ld\t0\t0\t0\t0\t0\t0\t%c\t%s\t%zu\t%ld\t%ld\t%s\t%zu\t%ld\t%ld\t1\t%ld,\t%ld,\t%ld,\n", m, (qe - qa) - m, strand ? '-' : '+',
                                     qs.names[q].c_str(), s.size(), qa, qe, db.names[t].c_str(), (size_t)tl, qa + diag, qe + diag, qe - qa, qa, qa + diag);
                    out.put(line, (size_t)n);
                }
            }
        }
}

// =============================================================================================================================
// removeMisassembly (AG:3821-4297)
// =============================================================================================================================
void ag_coverage_pileup_host(const std::string& sam_path, const std::vector<u32>& chunk_len, std::vector<int>& cov, void*) {   // AG:3923-3970
    std::vector<size_t> off(chunk_len.size() + 1, 0);
    for (size_t i = 0; i < chunk_len.size(); i++) off[i + 1] = off[i] + chunk_len[i];
    cov.assign(off.back(), 0);
    FileMap fm(sam_path);
    if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
    Lines ln(fm.p, fm.n); const char* s; size_t n;
    std::vector<ag_seg> s1, s2;
    auto add = [&](const SamRec& r) {
        if (r.tid >= chunk_len.size()) return;   // (the reference indexes without a check)
        for (u32 bp = r.tstart; bp < r.tend && bp < chunk_len[r.tid]; bp++) cov[off[r.tid] + bp]++;
    };
    while (ln.next(s, n)) {
        if (n == 0 || s[0] == 0) break;
        if (s[0] == '@') continue;
        SamRec a, b;
        s1.clear(); s2.clear();
        parse_sam(s, n, a, s1);
        if (!ln.next(s, n) || n == 0 || s[0] == 0) throw AgHostError{"BROKEN BOWTIE FILE!"};
        parse_sam(s, n, b, s2);
        if (a.tid != AG_NONE && b.tid != AG_NONE) { add(a); add(b); }
    }
}

namespace {
struct MPos { u32 tid, ss, se, ts, te, fr; };   // ContigPosition; tid == NONE: deleted entry (p0)
const MPos kP0 = {AG_NONE, AG_NONE, AG_NONE, AG_NONE, AG_NONE, AG_NONE};
inline int m_conflict(u32 x1, u32 y1, u32 x2, u32 y2) {   // AG:3980-3988
    return ((x1 <= x2 && x2 <= y1 && y1 <= y2 && (int)y1 - (int)x2 >= 100) || (x2 <= x1 && x1 <= y2 && y2 <= y1 && (int)y2 - (int)x1 >= 100) ||
            (x1 <= x2 && x2 <= y2 && y2 <= y1 && (int)y2 - (int)x2 >= 100) || (x2 <= x1 && x1 <= y1 && y1 <= y2 && (int)y1 - (int)x1 >= 100) ||
            (x1 <= x2 && y2 <= y1) || (x2 <= x1 && y1 <= y2)) ? 1 : 0;
}
inline int m_close(u32 y1, u32 x2, u32 threshold) { return (u32)abs((int)x2 - (int)y1) < threshold ? 1 : 0; }   // AG:3990-3996 (int < unsigned compares as unsigned)
}  // namespace

void ag_remove_misassembly(const std::string& file, const std::string& id, int coverage, const std::string& tmp, AgAlignFn align, AgPileupFn pileup, void* user) {
    const double kMinThreshold = 0.1;   // MIN_THRESHOLD, AG:42
    const std::string cfa = tmp + "/_" + id + "_contigs.fa";
    std::vector<std::string> contig_ids;
    ag_formalize_contigs_to(file, cfa, "", contig_ids);                         // formalizeInput(c, "tmp/_<id>_contigs.fa") — also resets contigIds
    if (!align(id, user)) throw AgHostError{"BLAT CALL FAILED!"};               // makeAlignment, AG:3821-3850
    // ---- loadPreContigs (AG:3852-3890): the chunks; loadContigs (AG:3892-3936): chunks regrouped into contigs ----
    std::vector<std::string> chunk; std::vector<int> chunk_real;
    {
        FileMap fm(cfa);
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            if (s[0] == '>') { chunk.emplace_back(); const char* dot = (const char*)memchr(s, '.', n); chunk_real.push_back(dot ? ag_atoi(dot + 1, std::min<size_t>((size_t)(s + n - dot - 1), 9)) : ag_atoi(s + n, 0)); }
            else if (!chunk.empty()) chunk.back().append(s, n);
        }
    }
    std::vector<u32> chunk_len; for (auto& c : chunk) chunk_len.push_back((u32)c.size());
    std::vector<int> chunk_cov;
    pileup(tmp + "/_reads_" + id + "_contigs.bowtie", chunk_len, chunk_cov, user);   // loadReadAlignment, AG:3938-3978
    std::vector<std::string> base; std::vector<std::vector<int>> cov;                  // per contig
    {
        int real_bak = -1; size_t o = 0;
        for (size_t c = 0; c < chunk.size(); c++) {
            if (chunk_real[c] > real_bak) { base.emplace_back(); cov.emplace_back(); real_bak = chunk_real[c]; }
            if (base.empty()) { o += chunk[c].size(); continue; }   // (the reference would write through an empty vector)
            base.back() += chunk[c];
            cov.back().insert(cov.back().end(), chunk_cov.begin() + (long)o, chunk_cov.begin() + (long)(o + chunk[c].size()));
            o += chunk[c].size();
        }
    }
    // ---- loadContigAlignment (AG:4003-4145) ----
    std::vector<std::vector<MPos>> positions(base.size());
    {
        FileMap fm(tmp + "/_" + id + "_contigs_genome.psl");
        if (!fm.ok) throw AgHostError{"CANNOT OPEN FILE!"};
        Lines ln(fm.p, fm.n); const char* s; size_t n;
        PslRec r; int real_bak = -1; u32 source_bak = 0;
        while (ln.next(s, n)) {
            if (n == 0 || s[0] == 0) break;
            parse_psl(s, n, r);
            // parseBLAT's return value: the id after the '.' of the query name, else the query size (AG:505-521)
            int real = (int)r.ssize;
            { const char* f = s; int tab = 0; const char* fe = s + n;
              for (const char* c = s; c <= fe; c++) if (c == fe || *c == '\t') { if (tab == 9) { const char* dot = (const char*)memchr(f, '.', (size_t)(c - f)); if (dot) real = ag_atoi(dot + 1, (size_t)(c - dot - 1)); } tab++; f = c + 1; } }
            if (real > real_bak) { real_bak = real; source_bak = r.sid; }
            u32 ss = (r.sid - source_bak) * (u32)kLargeChunk + r.sstart, se = (r.sid - source_bak) * (u32)kLargeChunk + r.send;
            if (!(se - ss >= 100 && (double)(se - ss - r.sgap) / (se - ss) >= kMinThreshold && (double)(r.tend - r.tstart - r.tgap) / (double)(r.tend - r.tstart) >= kMinThreshold)) continue;
            if (real < 0 || (size_t)real >= positions.size()) continue;   // (out of range in the reference)
            std::vector<MPos>& ps = positions[(size_t)real];
            int keep = 1;
            for (size_t pp = 0; pp < ps.size(); pp++)
                if (ps[pp].tid != AG_NONE && r.tid == ps[pp].tid && m_conflict(ss, se, ps[pp].ss, ps[pp].se)) {
                    if (se - ss < ps[pp].se - ps[pp].ss) keep = 0; else ps[pp] = kP0;
                }
            if (keep) ps.push_back(MPos{r.tid, ss, se, r.tstart, r.tend, r.fr});
        }
    }
    for (auto& ps : positions)   // join local alignments that continue each other (AG:4068-4082); the restart quirk (ppp = 0, then ++) is kept
        for (size_t pp = 0; pp < ps.size(); pp++)
            for (size_t ppp = 0; ppp < ps.size(); ppp++)
                if (ppp != pp && ps[pp].tid != AG_NONE && ps[ppp].tid != AG_NONE && ps[pp].tid == ps[ppp].tid &&
                    m_close(ps[pp].se, ps[ppp].ss, (u32)(abs((int)ps[pp].se - (int)ps[pp].ss) / 10)) && m_close(ps[pp].te, ps[ppp].ts, (u32)(abs((int)ps[pp].te - (int)ps[pp].ts) / 10)) &&
                    ps[pp].fr == ps[ppp].fr) {
                    ps[pp].se = ps[ppp].se; ps[pp].te = ps[ppp].te; ps[ppp] = kP0; ppp = 0;
                }
    for (auto& ps : positions)   // conflicting alignments to different places: keep the longer (AG:4084-4091)
        for (size_t pp = 0; pp < ps.size(); pp++)
            for (size_t ppp = pp + 1; ppp < ps.size(); ppp++)
                if (ps[pp].tid != AG_NONE && ps[ppp].tid != AG_NONE && m_conflict(ps[pp].ss, ps[pp].se, ps[ppp].ss, ps[ppp].se)) {
                    if (ps[pp].se - ps[pp].ss > ps[ppp].se - ps[ppp].ss) ps[ppp] = kP0; else ps[pp] = kP0;
                }
    for (size_t sp = 0; sp < positions.size(); sp++) {   // keep a distance between consecutive local alignments: cut at the coverage minimum (AG:4094-4142)
        std::vector<MPos>& ps = positions[sp]; const std::vector<int>& cv = cov[sp];
        auto cov_at = [&](long bp) -> int { return bp >= 0 && (size_t)bp < cv.size() ? cv[(size_t)bp] : 0; };
        for (size_t pp = 0; pp < ps.size(); pp++)
            for (size_t ppp = pp + 1; ppp < ps.size(); ppp++) {
                if (ps[pp].tid == AG_NONE || ps[ppp].tid == AG_NONE) continue;
                if (overlap(ps[pp].ss, ps[pp].se, ps[ppp].ss, ps[ppp].se)) {
                    int mn = 99999, mp = -1, start, end;
                    if (ps[pp].ss <= ps[ppp].ss) { start = (int)ps[ppp].ss; end = (int)ps[pp].se - 1; } else { start = (int)ps[pp].ss; end = (int)ps[ppp].se - 1; }
                    for (int bp = start; bp <= end; bp++) if (cov_at(bp) < mn) { mn = cov_at(bp); mp = bp; }
                    if (ps[pp].ss <= ps[ppp].ss) { ps[pp].se = (u32)mp; ps[ppp].ss = (u32)(mp + 1); } else { ps[ppp].se = (u32)mp; ps[pp].ss = (u32)(mp + 1); }
                } else if (ps[pp].se == ps[ppp].ss) {
                    if (cov_at((long)ps[pp].se - 1) < cov_at((long)ps[ppp].ss)) ps[pp].se--; else ps[ppp].ss++;
                } else if (ps[ppp].se == ps[pp].ss) {
                    if (cov_at((long)ps[ppp].se - 1) < cov_at((long)ps[pp].ss)) ps[ppp].se--; else ps[pp].ss++;
                }
            }
    }
    // ---- removeMasb (AG:4147-4279): -1 = safe, -2 = removed ----
    for (size_t sp = 0; sp < positions.size(); sp++) {
        std::vector<MPos>& ps = positions[sp]; std::vector<int>& cv = cov[sp];
        bool whole = false;
        for (const MPos& p : ps) if (p.tid != AG_NONE && (double)(p.se - p.ss) / cv.size() >= 0.8) { whole = true; break; }
        if (whole) { std::fill(cv.begin(), cv.end(), -1); continue; }
        for (const MPos& p : ps) if (p.tid != AG_NONE) for (u32 bp = p.ss; bp < p.se && bp < cv.size(); bp++) cv[bp] = -1;
        const long sz = (long)cv.size();
        int start = 0, end = 0, total = 0;
        for (long bp = 0; bp < sz; bp++) {
            if (cv[bp] == -1) continue;
            if (bp != 0 && bp != sz - 1 && cv[bp - 1] == -1 && cv[bp + 1] == -1) { cv[bp] = cv[bp] < coverage ? -2 : -1; continue; }
            if (bp == 0 || cv[bp - 1] == -1) { start = (int)bp; total = cv[bp]; }
            else if (bp == sz - 1 || cv[bp + 1] == -1) {
                end = (int)bp; total += cv[bp];
                const int mark = total / (end - start + 1) < coverage ? -2 : -1;
                for (int b = start; b <= end; b++) cv[b] = mark;
            } else total += cv[bp];
        }
    }
    Out out(std::string("corrected_") + file);
    for (size_t cp = 0; cp < base.size(); cp++) {
        const std::vector<int>& cv = cov[cp]; const std::string& bs = base[cp];
        std::vector<std::string> split;
        const long sz = (long)cv.size();
        for (long bp = 0; bp < sz; bp++) {
            if ((split.empty() && cv[bp] == -1) || (bp > 0 && cv[bp - 1] == -2 && cv[bp] == -1)) split.emplace_back();
            if (cv[bp] == -1 && !split.empty()) split.back().push_back(bs[(size_t)bp]);
            if (bp == sz - 1 || (cv[bp] == -1 && cv[bp + 1] == -2))
                if (!split.empty() && split.back().size() <= 200) split.pop_back();
        }
        for (size_t sp = 0; sp < split.size(); sp++) {
            out.ch('>'); if (cp < contig_ids.size()) out.put(contig_ids[cp]);
            if (split.size() != 1) { out.put(" : part", 7); out.num(sp); }
            out.ch('\n'); out.wrap60(split[sp]);
        }
    }
    if (id == "remaining") {
        FileMap fm(tmp + "/_chaff.fa");
        if (fm.ok) { Lines ln(fm.p, fm.n); const char* s; size_t n; while (ln.next(s, n)) { if (n == 0 || s[0] == 0) break; out.put(s, n); out.ch('\n'); } }
    }
}

// The host side stages hundreds of MB per unit in vectors.  By default glibc serves such blocks with mmap and returns them on free, so
// every unit pays the page faults (and the kernel's page zeroing) again; keeping them on the heap makes the second and later units run
// on warm memory.  This is process-wide allocator policy, so the library only applies it when asked to (AG_MALLOPT=1, see ag_create); the
// drop-in CLI applies it in its own main().  AG_NO_MALLOPT=1 switches it off everywhere.
void ag_tune_malloc() {
    static bool done = false;
    if (done || getenv("AG_NO_MALLOPT")) return;
    done = true;
    mallopt(M_MMAP_MAX, 0);
    mallopt(M_TRIM_THRESHOLD, 0x7FFFFFFF);
    mallopt(M_TOP_PAD, 64 << 20);
}

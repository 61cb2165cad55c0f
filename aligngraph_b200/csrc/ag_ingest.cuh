// GPU-side ingestion of the hot path's text inputs: the raw bytes of tmp/_reads.fa (loadSeq, AG:361-404) and of
// tmp/_reads_genome.N.bowtie (parseBOWTIE AG:181-285 + loadReadAli AG:1233-1277 + the duplicate / strand rules AG:1650-1671) are copied to
// the device through page-locked staging chunks and parsed there; the resulting packed reads and `ag_aln` tuples never visit the host.
// Included by ag_device.cu (shares its buffers, scanner and error macros).
//
//   k_nl_count / k_nl_fill    newline index of a text buffer: 64 bytes per thread (4 x uint4 loads), count -> scan -> positions
//   k_rd_len / k_rd_pack      reads: record = '>' line + one sequence line; length per record, then 32 bases per thread -> 2-bit words,
//                             non-ACGT bit plane, exception list (read * 65536 + offset, original character)
//   k_sam_parse               one thread per record PAIR through the allocation-free functions of ag_samcore.h (field split, CIGAR state
//                             machine, the `double` ratio filter of AG:1261, segment normalisation)
//   k_sam_sorted / k_sam_lost / k_sam_survive / k_sam_fill
//                             the order-dependent half in data-parallel form: ids non-decreasing; the records lost at the 1,000,000-id batch
//                             boundaries (AG:1259) by binary search; the duplicate rule (AG:1650-1655) by looking back through the record's
//                             group; count -> scan -> fill of the surviving `ag_aln` tuples in file order
// Anything outside the well-formed layout (multi-line reads, empty lines, unsorted ids, unknown CIGAR characters, ...) is reported as
// "not well formed" and the sequential host parser — the literal semantics, error messages included — takes the file.
#pragma once
#include "ag_samcore.h"

namespace {

// ---------------------------------------------------------------------------------------------------------------------------
// newline index
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int NL_T = 256, NL_BYTES = 64, NL_B = NL_T * NL_BYTES;   // 16 KB of text per CTA

__device__ __forceinline__ u32 nl_bits(u32 w) { return __vcmpeq4(w, 0x0A0A0A0Au) & 0x01010101u; }   // bit 0 of every byte that is '\n'

__global__ void __launch_bounds__(NL_T) k_nl_count(const uint4* __restrict__ text, size_t n16, u32* __restrict__ blk) {
    __shared__ u32 sm[33];
    const size_t i0 = ((size_t)blockIdx.x * NL_T + threadIdx.x) * (NL_BYTES / 16);
    u32 c = 0;
#pragma unroll
    for (int j = 0; j < NL_BYTES / 16; j++)
        if (i0 + j < n16) { const uint4 v = text[i0 + j]; c += __popc(nl_bits(v.x)) + __popc(nl_bits(v.y)) + __popc(nl_bits(v.z)) + __popc(nl_bits(v.w)); }
    u32 total; block_excl_scan(c, sm, total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}
__global__ void __launch_bounds__(NL_T) k_nl_fill(const uint4* __restrict__ text, size_t n16, const u32* __restrict__ blk_off, u32* __restrict__ nl) {
    __shared__ u32 sm[33];
    const size_t i0 = ((size_t)blockIdx.x * NL_T + threadIdx.x) * (NL_BYTES / 16);
    uint4 v[NL_BYTES / 16]; u32 c = 0;
#pragma unroll
    for (int j = 0; j < NL_BYTES / 16; j++) {
        v[j] = make_uint4(0, 0, 0, 0);
        if (i0 + j < n16) { v[j] = text[i0 + j]; c += __popc(nl_bits(v[j].x)) + __popc(nl_bits(v[j].y)) + __popc(nl_bits(v[j].z)) + __popc(nl_bits(v[j].w)); }
    }
    u32 total; u32 o = block_excl_scan(c, sm, total) + blk_off[blockIdx.x];
    if (!c) return;
#pragma unroll
    for (int j = 0; j < NL_BYTES / 16; j++) {
        const u32 w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 b = nl_bits(w[k]);
            while (b) { const u32 byte = (u32)(__ffs((int)b) - 1) >> 3; b &= b - 1; nl[o++] = (u32)((i0 + j) * 16 + k * 4 + byte); }
        }
    }
}

// line `i` of a text whose newline positions are nl[]: [start, end) without the newline; `body` = offset of line 0
__device__ __forceinline__ void line_span(const u32* __restrict__ nl, u32 body, u32 i, u32& s, u32& e) { s = i ? nl[i - 1] + 1 : body; e = nl[i]; }

// ---------------------------------------------------------------------------------------------------------------------------
// reads
// ---------------------------------------------------------------------------------------------------------------------------
enum { ING_BAD_LAYOUT = 1, ING_BAD_RECORD = 2, ING_BAD_ORDER = 8, ING_STRAND = 32, ING_PE_LEN = 64, ING_TOO_LONG = 128, ING_WINDOW = 256 };

// record r = lines 2r ('>' header) and 2r + 1 (sequence); rlen[r] = bases; flags: layout errors, maximum length
__global__ void k_rd_len(const char* __restrict__ text, const u32* __restrict__ nl, u32 n_rec, u32* __restrict__ rlen, u32* maxlen, int* bad) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    u32 hs, he, ss, se;
    line_span(nl, 0, 2 * r, hs, he); line_span(nl, 0, 2 * r + 1, ss, se);
    if (text[hs] != '>' || se == ss || text[ss] == '>' || text[ss] == 0 || he == hs) atomicOr(bad, ING_BAD_LAYOUT);   // multi-line / empty records: sequential host parser
    const u32 len = se - ss;
    rlen[r] = len;
    if (len > 65535u) atomicOr(bad, ING_TOO_LONG);
    u32 m = len;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxlen, m);
}
__global__ void k_rd_pairlen(const u32* __restrict__ rlen, u32 n_pairs, uint16_t* __restrict__ pair_len, int* bad) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const u32 a = rlen[2 * p], b = rlen[2 * p + 1];
    if (a != b) atomicOr(bad, ING_PE_LEN);   // INCONSISTENT PE FILES!
    pair_len[p] = (uint16_t)a;
}
// one thread per (record, group of 32 bases): two 2-bit words + one mask word; characters other than upper-case ACGT are masked and listed.
// The 32 characters are fetched as aligned 32-bit words (funnel-shifted to the sequence's byte offset) and coded four at a time: for A C G T
// ((x >> 1) ^ (x >> 2)) & 3 is the 2-bit code, and rebuilding the character from the code tells whether the byte was one of the four.
__global__ void k_rd_pack(const char* __restrict__ text, const u32* __restrict__ nl, u32 n_rec, u32 stride2, u32 stridem, u64 read0,
                          u32* __restrict__ bases, u32* __restrict__ nmask, u64* __restrict__ exc_key, char* __restrict__ exc_chr, u32* exc_count, u32 exc_cap) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 r = (u32)(t / stridem), g = (u32)(t % stridem);
    if (r >= n_rec) return;
    const u32 ss = nl[2 * r] + 1, se = nl[2 * r + 1];
    const u32 len = se - ss, o0 = g * 32;
    u32 w0 = 0, w1 = 0, mk = 0;
    if (o0 < len) {
        const u32 cnt = min(32u, len - o0);
        const char* s = text + ss + o0;
        const unsigned long long addr = (unsigned long long)s;
        const u32* wp = reinterpret_cast<const u32*>(addr & ~3ull);   // (the text buffer is 16-byte aligned and padded: the words around the sequence are readable)
        const u32 sh = (u32)(addr & 3ull) * 8u;
        u32 prev = wp[0];
#pragma unroll
        for (u32 k = 0; k < 8; k++) {
            if (4 * k >= cnt) break;
            const u32 next = wp[k + 1];
            const u32 x = __funnelshift_r(prev, next, sh);   // characters 4k .. 4k+3
            prev = next;
            const u32 c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
            const u32 c0 = c & 0x01010101u, c1 = (c >> 1) & 0x01010101u;
            const u32 expect = 0x41414141u + c0 * 2u + c1 * 6u + (c0 & c1) * 0x0Bu;   // 'A' 'C' 'G' 'T' rebuilt from the code
            u32 diff = x ^ expect;                                                     // non-zero byte = not one of the four
            const u32 valid = min(4u, cnt - 4 * k);
            if (valid < 4) diff &= (1u << (8 * valid)) - 1u;
            u32 codes = c;
            if (diff) {
                for (u32 j = 0; j < valid; j++)
                    if ((diff >> (8 * j)) & 0xFFu) {
                        mk |= 1u << (4 * k + j);
                        codes &= ~(3u << (8 * j));
                        const u32 e = atomicAdd(exc_count, 1u);
                        if (e < exc_cap) { exc_key[e] = (read0 + r) * 65536ull + (o0 + 4 * k + j); exc_chr[e] = (char)((x >> (8 * j)) & 0xFFu); }
                    }
            }
            if (valid < 4) codes &= (1u << (8 * valid)) - 1u;
            const u32 packed = (codes & 3u) | ((codes >> 6) & 0xCu) | ((codes >> 12) & 0x30u) | ((codes >> 18) & 0xC0u);   // four 2-bit codes -> one byte
            if (k < 4) w0 |= packed << (8 * k); else w1 |= packed << (8 * (k - 4));
        }
    }
    const u64 rr = read0 + r;
    bases[rr * stride2 + 2 * g] = w0;
    if (2 * g + 1 < stride2) bases[rr * stride2 + 2 * g + 1] = w1;
    nmask[rr * stridem + g] = mk;
}

// ---------------------------------------------------------------------------------------------------------------------------
// SAM
// ---------------------------------------------------------------------------------------------------------------------------
// one record pair after the per-record half: what the order-dependent half and the fill need (32 bytes, the shape of ag_aln)
struct ag_srec { u32 sid, flags /* bit0 fr1, bit1 fr2, bit2 passes AG:1261, bits 8-15 n1, bits 16-23 n2 */, p0, dst1, sl1, dst2, sl2, next /* ext segments */; };

__device__ __forceinline__ bool sam_parse_pair(const char* __restrict__ text, const u32* __restrict__ nl, u32 body, u32 p, const uint16_t* __restrict__ pair_len, u64 n_read_pairs,
                                               u64 win_lo, u64 win_hi, ag_srec& r, ag_seg* n1, ag_seg* n2, int* bad) {
    u32 s0, e0, s1, e1;
    line_span(nl, body, 2 * p, s0, e0); line_span(nl, body, 2 * p + 1, s1, e1);
    r.sid = 0; r.flags = 0; r.p0 = AG_NONE; r.dst1 = r.sl1 = r.dst2 = r.sl2 = 0; r.next = 0;
    if (e0 == s0 || e1 == s1 || text[s0] == '@' || text[s1] == '@' || text[s0] == 0 || text[s1] == 0) { atomicOr(bad, ING_BAD_LAYOUT); return false; }
    ag_samline a, b;
    ag_sam_parse_line(text + s0, e0 - s0, a);
    ag_sam_parse_line(text + s1, e1 - s1, b);
    if (a.err || b.err) { atomicOr(bad, ING_BAD_RECORD); return false; }   // unknown CIGAR character / very long CIGAR: the host parser reports or handles it
    r.sid = a.sid; r.flags = a.fr | (b.fr << 1);
    r.p0 = a.tid == AG_NONE ? AG_NONE : ag_sam_pos_at0(a.seg, a.nseg);
    if (ag_sam_mate_pass(a, 0.6) && ag_sam_mate_pass(b, 0.6)) {
        if (a.tid != 0 || b.tid != 0 || b.sid != a.sid || a.sid >= n_read_pairs) { atomicOr(bad, ING_BAD_RECORD); return false; }
        if (a.sid < win_lo || a.sid > win_hi) { atomicOr(bad, ING_WINDOW); return false; }   // this read is not resident (windowed read set): the caller loads the whole set
        const u32 rlen = pair_len[a.sid];
        u32 c1 = 0, c2 = 0;
        if (ag_sam_normalize(a.seg, a.nseg, rlen, n1, c1) || ag_sam_normalize(b.seg, b.nseg, rlen, n2, c2) || !c1 || !c2 || c1 > 255 || c2 > 255) { atomicOr(bad, ING_BAD_RECORD); return false; }
        r.flags |= 4u | (c1 << 8) | (c2 << 16);
        r.dst1 = n1[0].dst; r.sl1 = n1[0].src | (n1[0].len << 16);
        r.dst2 = n2[0].dst; r.sl2 = n2[0].src | (n2[0].len << 16);
        r.next = (c1 > 1 ? c1 : 0) + (c2 > 1 ? c2 : 0);
    }
    return true;
}

__global__ void __launch_bounds__(128) k_sam_parse(const char* __restrict__ text, const u32* __restrict__ nl, u32 body, u32 n_rec, const uint16_t* __restrict__ pair_len, u64 n_read_pairs,
                                                   u64 win_lo, u64 win_hi, ag_srec* __restrict__ rec, int* bad) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rec) return;
    ag_srec r; ag_seg n1[AG_SAM_MAXSEG], n2[AG_SAM_MAXSEG];
    sam_parse_pair(text, nl, body, p, pair_len, n_read_pairs, win_lo, win_hi, r, n1, n2, bad);
    rec[p] = r;
}
__global__ void k_sam_sorted(const ag_srec* __restrict__ rec, u32 n_rec, int* bad) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n_rec) return;
    if (rec[i + 1].sid < rec[i].sid) atomicOr(bad, ING_BAD_ORDER);
}
// Batches of 1,000,000 read ids (AG:1885-1894): the first record whose id lies beyond the current batch is consumed and LOST (AG:1259)
// and the batch advances by one; once the batch reaches the end of the read set the reference stops reading (`stop`).  ids are sorted.
// out: lost[0] = count, lost[1] = stop, lost[2..] = the lost record indices (ascending)
__global__ void k_sam_lost(const ag_srec* __restrict__ rec, u32 n_rec, long long n_read_pairs, u32* lost, u32 lost_cap) {
    if (blockIdx.x || threadIdx.x) return;
    const long long BATCH = 1000000;
    long long last = min(BATCH - 1, n_read_pairs - 1);
    u32 cnt = 0, stop = n_rec, i = 0;
    for (;;) {
        u32 lo = i, hi = n_rec;   // first index >= i whose id is > last
        while (lo < hi) { const u32 mid = lo + (hi - lo) / 2; if ((long long)rec[mid].sid > last) hi = mid; else lo = mid + 1; }
        const u32 j = lo;
        if (j >= n_rec) break;
        if (last >= n_read_pairs - 1) { stop = j; break; }
        if (cnt < lost_cap) lost[2 + cnt] = j;
        cnt++;
        last = min(last + BATCH, n_read_pairs - 1);
        i = j + 1;
    }
    lost[0] = cnt; lost[1] = stop;
}
__device__ __forceinline__ bool sam_is_lost(const u32* __restrict__ lost, u32 n_lost, u32 g) {
    u32 lo = 0, hi = n_lost;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; const u32 v = lost[2 + mid]; if (v == g) return true; if (v < g) lo = mid + 1; else hi = mid; }
    return false;
}
// keep[i] = 1 when record i passes AG:1261, is neither lost nor beyond `stop`, and no earlier record of its group (consecutive records that
// passed the filter with the same id, not separated by a lost record) lies within one read length (AG:1650-1655)
__global__ void k_sam_survive(const ag_srec* __restrict__ rec, u32 n_rec, const u32* __restrict__ lost, const uint16_t* __restrict__ pair_len, u32* __restrict__ keep, u32* __restrict__ next, int* bad) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const ag_srec r = rec[i];
    const u32 n_lost = lost[0], stop = lost[1];
    bool k = (r.flags & 4u) && i < stop && !(n_lost && sam_is_lost(lost, n_lost, i));
    if (k) {
        u32 rlen = 0; bool have = false;
        for (u32 j = i; j-- > 0;) {
            if (n_lost && sam_is_lost(lost, n_lost, j)) break;
            const ag_srec o = rec[j];
            if (!(o.flags & 4u)) continue;
            if (o.sid != r.sid) break;
            if (!have) { rlen = pair_len[r.sid]; have = true; }
            if (ag_absdiff(r.p0, o.p0) < (int)rlen) { k = false; break; }
        }
    }
    if (k && ((r.flags & 1u) == ((r.flags >> 1) & 1u))) atomicOr(bad, ING_STRAND);   // exactly one mate must be reverse (AG:1657-1671)
    keep[i] = k ? 1u : 0u;
    next[i] = k ? r.next : 0u;
}
__global__ void __launch_bounds__(128) k_sam_fill(const char* __restrict__ text, const u32* __restrict__ nl, u32 body, u32 n_rec, const uint16_t* __restrict__ pair_len, u64 n_read_pairs,
                                                  u64 win_lo, u64 win_hi, const ag_srec* __restrict__ rec, const u32* __restrict__ keep, const u32* __restrict__ aoff, const u32* __restrict__ eoff,
                                                  ag_aln* __restrict__ aln, ag_seg* __restrict__ ext) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec || !keep[i]) return;
    const ag_srec r = rec[i];
    ag_aln a; a.pair = r.sid; a.pad = 0; a.flags = r.flags & ~4u;
    a.dst1 = r.dst1; a.sl1 = r.sl1; a.dst2 = r.dst2; a.sl2 = r.sl2;
    u32 oe = eoff ? eoff[i] : 0u;
    a.ext_idx = oe;
    if (r.next) {   // multi-segment CIGAR (rare): parse the pair again and write the normalised segment lists
        ag_srec r2; ag_seg n1[AG_SAM_MAXSEG], n2[AG_SAM_MAXSEG]; int dummy = 0;
        sam_parse_pair(text, nl, body, i, pair_len, n_read_pairs, win_lo, win_hi, r2, n1, n2, &dummy);
        const u32 c1 = (r.flags >> 8) & 0xFF, c2 = (r.flags >> 16) & 0xFF;
        if (c1 > 1) { for (u32 j = 0; j < c1; j++) ext[oe + j] = n1[j]; oe += c1; }
        if (c2 > 1) { for (u32 j = 0; j < c2; j++) ext[oe + j] = n2[j]; }
    }
    aln[aoff[i]] = a;
}

// ---------------------------------------------------------------------------------------------------------------------------
// read-vs-contig coverage pile-up of removeMisassembly (loadReadAlignment overload, AG:3938-3978): every pair with both mates aligned adds
// one to the coverage of target bases [targetStart, targetEnd) of both mates.  One thread per record pair marks interval starts / ends in a
// difference array (atomics); an inclusive scan turns it into the per-base coverage.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_cov_marks(const char* __restrict__ text, const u32* __restrict__ nl, u32 body, u32 n_rec, const u64* __restrict__ chunk_off, u32 n_chunks,
                                                   u32* __restrict__ diff, int* bad) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rec) return;
    u32 s0, e0, s1, e1;
    line_span(nl, body, 2 * p, s0, e0); line_span(nl, body, 2 * p + 1, s1, e1);
    if (e0 == s0 || e1 == s1 || text[s0] == '@' || text[s1] == '@' || text[s0] == 0 || text[s1] == 0) { atomicOr(bad, ING_BAD_LAYOUT); return; }
    ag_samline a, b;
    ag_sam_parse_line(text + s0, e0 - s0, a);
    ag_sam_parse_line(text + s1, e1 - s1, b);
    if (a.err == AG_SAM_ERR_CHAR || b.err == AG_SAM_ERR_CHAR) { atomicOr(bad, ING_BAD_RECORD); return; }   // `unknown character`: the host path reports it
    if (a.tid == AG_NONE || b.tid == AG_NONE) return;
    const ag_samline* r[2] = {&a, &b};
    for (int k = 0; k < 2; k++) {
        if (r[k]->tid >= n_chunks) continue;
        const u64 o = chunk_off[r[k]->tid], len = chunk_off[r[k]->tid + 1] - o;
        const u64 ts = r[k]->tstart, te = r[k]->tend < len ? r[k]->tend : len;
        if (ts >= te) continue;
        atomicAdd(&diff[o + ts], 1u);
        atomicAdd(&diff[o + te], 0xFFFFFFFFu);   // -1 (diff has one entry more than there are bases)
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// built-in containment search of refinement() (ag_contain_search): matching bases of every candidate placement of a query (either strand)
// on a database sequence.  One warp per candidate, lanes stride over the query (coalesced byte loads), shuffle reduction.
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_place { u32 q, strand, t, pad; long long start; };
__global__ void k_verify_placements(const char* __restrict__ db, const u64* __restrict__ db_off, const char* __restrict__ qs, const u64* __restrict__ q_off,
                                    const ag_place* __restrict__ cand, u32 n_cand, u32* __restrict__ match) {
    const u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_cand) return;
    const ag_place p = cand[w];
    const char* q = qs + q_off[p.q]; const u32 len = (u32)(q_off[p.q + 1] - q_off[p.q]);
    const char* t = db + db_off[p.t] + p.start;
    u32 m = 0;
    for (u32 i = lane; i < len; i += 32) {
        char c = p.strand ? q[len - 1 - i] : q[i];
        if (p.strand) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
        m += c == t[i];
    }
    for (int o = 16; o; o >>= 1) m += __shfl_xor_sync(0xFFFFFFFFu, m, o);
    if (lane == 0) match[w] = m;
}

// ---------------------------------------------------------------------------------------------------------------------------
// file -> device through a ring of page-locked PIECES (256 KB; 256 MB ring).  Reader threads draw pieces in file order, pread() them into their ring slots
// (page cache -> pinned memory) and raise the piece's ready flag; ONE submitter walks the pieces in order and queues a host->device copy for
// every run of consecutive ready pieces (up to 8 MB, not across the ring's wrap), so the copy engine starts after the first piece, not
// after every thread's first chunk, and the driver sees one caller and few, large copies.  Ring positions continue across calls (the reads
// file and the SAM file follow each other without waiting for the earlier copies); a slot is reused once the copy batch that read it has
// finished (one event per batch).
// ---------------------------------------------------------------------------------------------------------------------------
struct FileStager {
    static constexpr int MAXT = 32;
    size_t PIECE = (size_t)256 << 10, NP = 1024, MAXRUN = 32;   // AG_STAGE_PIECE_KB / AG_STAGE_RING_MB / AG_STAGE_RUN override (tuning)
    PinnedBuf ring;
    std::vector<cudaEvent_t> bev;              // event of batch b = bev[b % NP]
    std::vector<unsigned long long> slot_batch;   // batch whose completion frees the slot (0 = never used)
    unsigned long long seq = 0, batch = 0;       // global piece / batch counters (continue across calls)
    FileStager() {
        if (const char* e = getenv("AG_STAGE_PIECE_KB")) { const long v = atol(e); if (v >= 64 && v <= 65536) PIECE = (size_t)v << 10; }
        if (const char* e = getenv("AG_STAGE_RING_MB")) { const long v = atol(e); if (v >= 8 && v <= 4096) NP = std::max<size_t>(8, ((size_t)v << 20) / PIECE); }
        if (const char* e = getenv("AG_STAGE_RUN")) { const long v = atol(e); if (v >= 1 && v <= 64) MAXRUN = (size_t)v; }
    }
    void release() { ring.release(); for (cudaEvent_t e : bev) cudaEventDestroy(e); bev.clear(); slot_batch.clear(); seq = batch = 0; }
    // copies file bytes [off, off + len) to dst (device), asynchronously on `st` (every copy has been QUEUED when this returns)
    void run(int fd, size_t off, size_t len, char* dst, cudaStream_t st, int device) {
        if (!len) return;
        if (bev.empty()) { bev.resize(NP); slot_batch.assign(NP, 0); for (auto& e : bev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
        ring.ensure(NP * PIECE);
        const size_t n_pieces = (len + PIECE - 1) / PIECE;
        int tmax = MAXT; if (const char* e = getenv("AG_STAGE_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= MAXT) tmax = v; }
        const int T = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)ag_team_size(), (size_t)tmax), n_pieces + 1));
        static const bool no_read = getenv("AG_STAGE_NOREAD") != nullptr, no_copy = getenv("AG_STAGE_NOCOPY") != nullptr;   // diagnosis only: time the two halves apart
        const unsigned long long g0 = seq;
        std::atomic<int> failed(0);
        std::atomic<size_t> next_piece(0);
        std::atomic<unsigned long long> submitted(g0);          // every piece with a global number below this has its copy queued and slot_batch set
        std::unique_ptr<std::atomic<unsigned char>[]> ready(new std::atomic<unsigned char>[n_pieces]);
        for (size_t i = 0; i < n_pieces; i++) ready[i].store(0, std::memory_order_relaxed);
        auto relax = [](unsigned& spins) { if (++spins < 64) { __builtin_ia32_pause(); } else { std::this_thread::yield(); } };
        // slot of global piece g is free once the copy of piece g - NP (its previous occupant) has finished
        auto wait_slot = [&](unsigned long long g) -> bool {
            const size_t slot = (size_t)(g % NP);
            if (g >= NP) { unsigned spins = 0; while (submitted.load(std::memory_order_acquire) + NP <= g) { if (failed.load()) return false; relax(spins); } }
            const unsigned long long b = slot_batch[slot];
            if (b && cudaEventSynchronize(bev[(size_t)(b % NP)]) != cudaSuccess) return false;
            return true;
        };
        auto read_piece = [&](size_t k) -> bool {
            const unsigned long long g = g0 + k;
            if (!wait_slot(g)) return false;
            char* h = ring.p + (size_t)(g % NP) * PIECE;
            const size_t o = k * PIECE, n = std::min(PIECE, len - o);
            size_t a = 0;
            while (!no_read && a < n) { const ssize_t got = pread(fd, h + a, n - a, (off_t)(off + o + a)); if (got <= 0) return false; a += (size_t)got; }
            return true;
        };
        auto submit_run = [&](size_t k, size_t cnt) -> bool {     // pieces [k, k + cnt): consecutive in the file and in the ring
            const unsigned long long g = g0 + k;
            const size_t o = k * PIECE, n = std::min(cnt * PIECE, len - o);
            if (!no_copy && cudaMemcpyAsync(dst + o, ring.p + (size_t)(g % NP) * PIECE, n, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
            const unsigned long long b = ++batch;
            if (cudaEventRecord(bev[(size_t)(b % NP)], st) != cudaSuccess) return false;
            for (size_t i = 0; i < cnt; i++) slot_batch[(size_t)((g + i) % NP)] = b;
            submitted.store(g + cnt, std::memory_order_release);
            return true;
        };
        if (T == 1) {
            for (size_t k = 0; k < n_pieces; k++) { if (!read_piece(k)) { failed = 1; break; } if (!submit_run(k, 1)) { failed = 2; break; } }
        } else {
            ag_parallel_chunks(T, [&](int t) {
                cudaSetDevice(device);
                if (t == 0) {                                       // the submitter
                    size_t k = 0;
                    while (k < n_pieces) {
                        unsigned spins = 0;
                        while (!ready[k].load(std::memory_order_acquire)) { if (failed.load()) return; relax(spins); }
                        size_t cnt = 1;
                        const size_t slot = (size_t)((g0 + k) % NP);
                        while (cnt < MAXRUN && k + cnt < n_pieces && slot + cnt < NP && ready[k + cnt].load(std::memory_order_acquire)) cnt++;
                        if (!submit_run(k, cnt)) { failed = 2; return; }
                        k += cnt;
                    }
                    return;
                }
                for (;;) {                                          // readers
                    const size_t k = next_piece.fetch_add(1);
                    if (k >= n_pieces || failed.load()) return;
                    if (!read_piece(k)) { failed = 1; return; }
                    ready[k].store(1, std::memory_order_release);
                }
            });
        }
        seq = g0 + n_pieces;
        if (failed) throw AgError{failed == 1 ? "CANNOT OPEN FILE!" : "host->device copy of a staged chunk failed"};
    }
};

}  // namespace

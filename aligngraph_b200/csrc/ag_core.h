// Per-thread logic of the B200 AlignGraph kernels, written so that the SAME functions compile for the device (nvcc) and for a
// host-side emulation used only by the CPU test-suite (tests/emul).  The product always runs them inside the CUDA kernels of
// ag_device.cu; nothing here is a CPU fallback.
//
// Formulation.  The reference mutates `vector<KMer>` lists in read order (AG:1353-1624).  A call
// updateKMer(P, nextP, ...) touches only the lists at P and nextP and reads otherwise static data, so the node list of a
// position is a pure function of the ORDERED sequence of touches aimed at it.  We therefore give every unit position its own
// thread, feed it the alignments covering it in global alignment order, and let it run the reference's first-compatible
// clustering sequentially.  Two facts make this cheap:
//   (1) inside one alignment the "k2" touch of call i (create-if-absent at nextP, AG:1480-1587) is immediately followed by the
//       "k1" touch of call i+1 at the same position with identical fields (AG:1362-1477), so they fuse into one k1 touch; only
//       the k2 of an alignment's LAST call stands alone;
//   (2) founder fields never change and lists only grow at the end, so the item an event resolved to is simply the FIRST node
//       of the FINAL list compatible with the event's candidate — edges (AG:1590-1623) are rebuilt from the final table
//       without storing per-event items.
#pragma once
#include "ag_types.h"

#ifdef __CUDA_ARCH__
#define AG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define AG_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define AG_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define AG_ATOMIC_CAS(p, c, v) atomicCAS((p), (c), (v))
#define AG_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
template <class T> static inline T ag_host_add(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T ag_host_min(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T ag_host_max(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T ag_host_cas(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
#define AG_ATOMIC_ADD(p, v) ag_host_add((p), (v))
#define AG_ATOMIC_MIN(p, v) ag_host_min((p), (v))
#define AG_ATOMIC_MAX(p, v) ag_host_max((p), (v))
#define AG_ATOMIC_CAS(p, c, v) ag_host_cas((p), (c), (v))
#define AG_ATOMIC_OR(p, v) (*(p) |= (v))
#endif

// Tile geometry of the node / edge sweeps: a CTA of AG_TILE threads = 8 warps; a warp owns AG_WPOS = 31 consecutive unit positions
// (lanes 0-30) and its lane 31 replays the NEXT position as a halo, so that the item an alignment resolves to at q + 1 reaches the
// lane of q by one shuffle (the edge of the call that starts at q, AG:1590-1623, is then known inside the node sweep).  A tile owns
// AG_TPOS positions and additionally needs every alignment that touches its halo position.
#ifndef AG_TWARPS
#define AG_TWARPS 8          // warps per tile (a compile-time knob for tuning builds only: tools/build_variants.py)
#endif
#define AG_TILE (32 * AG_TWARPS)
#define AG_WPOS 31
#define AG_TPOS (AG_TWARPS * AG_WPOS)
AG_HD void ag_tile_range(u32 lo, u32 hi, u32 n_tiles, u32& t0, u32& t1) {  // tiles that must see an alignment touching [lo, hi]
    t0 = lo ? (lo - 1) / AG_TPOS : 0u;
    t1 = hi / AG_TPOS;
    if (t1 >= n_tiles) t1 = n_tiles - 1;
    if (t0 > t1) t0 = t1;
}

AG_HD u32 ag_min_u32(u32 a, u32 b) { return a < b ? a : b; }
AG_HD int ag_absdiff(u32 a, u32 b) {  // abs((int)(a - b)) on unsigned operands, as the reference writes it (AG:1296)
    int d = (int)(a - b);
    return d < 0 ? -d : d;
}

// ---------------------------------------------------------------------------------------------------------------------------
// segments
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_segv {  // a mate's segment list: inline single segment or a slice of the ext array
    u32 n, dst0, sl0;
    const ag_seg* ext;
    AG_HD ag_seg get(u32 j) const {
        if (n == 1) { ag_seg s; s.src = sl0 & 0xFFFFu; s.dst = dst0; s.len = sl0 >> 16; return s; }
        return ext[j];
    }
    // unit position of read offset `off`, or NONE (unaligned)
    AG_HD u32 pos_at(u32 off) const {
        for (u32 j = 0; j < n; j++) { ag_seg s = get(j); if (off - s.src < s.len) return s.dst + (off - s.src); }
        return AG_NONE;
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// k_prep: resolve left/right mate and the range of unit positions the alignment touches
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_prep_out { ag_alnp p; u32 lo, span; int any; };

// number of calls c (= number of fused k1 touches) of a left mate: aligned offsets a_0 < a_1 < ... ; call i exists iff
// a_i < L - k (loop bound AG:1681) and a_{i+1} exists (a successor is found, AG:1695-1700 / ordinary case).
AG_HD u32 ag_num_calls(const ag_segv& L, u32 len, u32 k) {
    u32 limit = len > k ? len - k : 0, m = 0, below = 0;
    for (u32 j = 0; j < L.n; j++) {
        ag_seg s = L.get(j);
        m += s.len;
        if (limit > s.src) below += (limit - s.src < s.len) ? (limit - s.src) : s.len;
    }
    if (m == 0) return 0;
    return below < m - 1 ? below : m - 1;
}

AG_HD ag_prep_out ag_prep(const ag_aln& a, const ag_seg* ext, u32 len, u32 k) {
    ag_prep_out o;
    u32 n1 = (a.flags >> 8) & 0xFF, n2 = (a.flags >> 16) & 0xFF;
    if (n1 == 1 && n2 == 1) {   // one M segment per mate (by far the common case): the loops below collapse to a few adds and compares
        const u32 src1 = a.sl1 & 0xFFFFu, len1 = a.sl1 >> 16, src2 = a.sl2 & 0xFFFFu, len2 = a.sl2 >> 16;
        const u32 limit = len > k ? len - k : 0;
        const u32 lo = src1 > src2 ? src1 : src2;
        u32 hi = ag_min_u32(src1 + len1, src2 + len2); if (hi > limit) hi = limit;
        const bool swap = lo < hi && (a.dst1 + (lo - src1)) > (a.dst2 + (lo - src2));   // AG:1672-1679
        const u32 lsrc = swap ? src2 : src1, llen = swap ? len2 : len1, ldst = swap ? a.dst2 : a.dst1;
        const u32 frL = swap ? ((a.flags >> 1) & 1) : (a.flags & 1);
        o.p.left_read = ((2 * a.pair + (swap ? 1 : 0)) << 1) | frL;
        o.p.len_nseg = len | (1u << 16) | (1u << 24);
        o.p.l_dst = ldst; o.p.l_sl = swap ? a.sl2 : a.sl1; o.p.r_dst = swap ? a.dst1 : a.dst2; o.p.r_sl = swap ? a.sl1 : a.sl2;
        o.p.ext_l = a.ext_idx; o.p.ext_r = a.ext_idx;
        const u32 below = limit > lsrc ? ag_min_u32(limit - lsrc, llen) : 0;
        const u32 c = llen == 0 ? 0 : ag_min_u32(below, llen - 1);                       // ag_num_calls
        o.any = c > 0; o.lo = c ? ldst : 0; o.span = c;
        return o;
    }
    ag_segv m1, m2;
    m1.n = n1; m1.dst0 = a.dst1; m1.sl0 = a.sl1; m1.ext = ext + a.ext_idx;
    m2.n = n2; m2.dst0 = a.dst2; m2.sl0 = a.sl2; m2.ext = ext + a.ext_idx + (n1 > 1 ? n1 : 0);
    // AG:1672-1679: mate 1 is the left mate unless, at some offset < L-k where both are aligned, it lies to the right
    u32 limit = len > k ? len - k : 0;
    bool swap = false;
    for (u32 i = 0; i < n1 && !swap; i++) {
        ag_seg s1 = m1.get(i);
        for (u32 j = 0; j < n2; j++) {
            ag_seg s2 = m2.get(j);
            u32 lo = s1.src > s2.src ? s1.src : s2.src;
            u32 hi1 = s1.src + s1.len, hi2 = s2.src + s2.len;
            u32 hi = hi1 < hi2 ? hi1 : hi2;
            if (hi > limit) hi = limit;
            if (lo < hi && (s1.dst + (lo - s1.src)) > (s2.dst + (lo - s2.src))) { swap = true; break; }
        }
    }
    const ag_segv& L = swap ? m2 : m1;
    const ag_segv& R = swap ? m1 : m2;
    u32 frL = swap ? ((a.flags >> 1) & 1) : (a.flags & 1);
    o.p.left_read = ((2 * a.pair + (swap ? 1 : 0)) << 1) | frL;
    o.p.len_nseg = len | (L.n << 16) | (R.n << 24);
    o.p.l_dst = L.dst0; o.p.l_sl = L.sl0; o.p.r_dst = R.dst0; o.p.r_sl = R.sl0;
    o.p.ext_l = (u32)(L.ext - ext); o.p.ext_r = (u32)(R.ext - ext);
    u32 c = ag_num_calls(L, len, k);
    o.any = c > 0;
    o.lo = 0; o.span = 0;
    if (c > 0) {
        o.lo = L.get(0).dst;
        u32 idx = c, hi = o.lo;  // position of aligned offset number c (the stand-alone k2)
        for (u32 j = 0; j < L.n; j++) { ag_seg s = L.get(j); if (idx < s.len) { hi = s.dst + idx; break; } idx -= s.len; }
        o.span = hi - o.lo;
    }
    return o;
}

AG_HD void ag_alnp_segs(const ag_alnp& p, const ag_seg* ext, ag_segv& L, ag_segv& R) {
    L.n = (p.len_nseg >> 16) & 0xFF; L.dst0 = p.l_dst; L.sl0 = p.l_sl; L.ext = ext + p.ext_l;
    R.n = (p.len_nseg >> 24) & 0xFF; R.dst0 = p.r_dst; R.sl0 = p.r_sl; R.ext = ext + p.ext_r;
}

// ---------------------------------------------------------------------------------------------------------------------------
// locate the touch an alignment makes at unit position q  (inverse of the offset loop AG:1681-1859)
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_touch {
    int kind;            // 0 none, 1 = k1 (a call starts here), 2 = stand-alone k2 (last call's successor)
    u32 soff, slen;      // founder string s = read[soff, soff+slen) in the oriented left mate; slen 0 = empty (gap chain)
    u32 mate;            // mate position (chromosomeOffset0) or NONE
    u32 npos, nmate, nsoff, nslen;  // successor of the call (kind 1 only)
};


AG_HD ag_touch ag_locate(const ag_alnp& p, const ag_seg* ext, u32 q, u32 k) {
    ag_touch t; t.kind = 0; t.soff = t.slen = 0; t.mate = AG_NONE; t.npos = t.nmate = AG_NONE; t.nsoff = t.nslen = 0;
    ag_segv L, R; ag_alnp_segs(p, ext, L, R);
    u32 len = p.len_nseg & 0xFFFFu;
    u32 c = ag_num_calls(L, len, k);
    if (c == 0) return t;
    u32 pre = 0;
    for (u32 j = 0; j < L.n; j++) {
        ag_seg s = L.get(j);
        if (q - s.dst < s.len) {  // q carries aligned offset a = number `i` in the aligned-offset order
            u32 d = q - s.dst, a = s.src + d, i = pre + d;
            if (i < c) {
                t.kind = 1; t.soff = a; t.slen = k; t.mate = R.pos_at(a);
                u32 a2, p2;  // next aligned offset and its position
                if (d + 1 < s.len) { a2 = a + 1; p2 = q + 1; }
                else { ag_seg s2 = L.get(j + 1); a2 = s2.src; p2 = s2.dst; }
                if (a2 == a + 1 || p2 == q + 1) {  // ordinary / deletion (AG:1791-1857) or pure insertion (AG:1707-1727)
                    t.npos = p2; t.nmate = R.pos_at(a2); t.nsoff = a2; t.nslen = ag_min_u32(k, len - a2);
                } else {                            // insertion followed by a deletion: gap chain (AG:1730-1750)
                    t.npos = q + 1; t.nmate = AG_NONE; t.nsoff = 0; t.nslen = 0;
                }
            } else if (i == c) {
                t.kind = 2; t.soff = a; t.slen = ag_min_u32(k, len - a); t.mate = R.pos_at(a);
            }
            return t;
        }
        if (j + 1 < L.n) {
            ag_seg s2 = L.get(j + 1);
            if (q >= s.dst + s.len && q < s2.dst) {  // inside the reference gap between two segments
                u32 last = pre + s.len - 1;
                if (last < c && s2.src > s.src + s.len) {  // the call from the segment's last base exists and read bases were skipped
                    t.kind = 1; t.soff = 0; t.slen = 0; t.mate = AG_NONE;
                    t.npos = q + 1;
                    if (q + 1 == s2.dst) { t.nmate = R.pos_at(s2.src); t.nsoff = s2.src; t.nslen = ag_min_u32(k, len - s2.src); }
                    else { t.nmate = AG_NONE; t.nsoff = 0; t.nslen = 0; }
                }
                return t;
            }
        }
        pre += s.len;
    }
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------------
// fast path: both mates are a single M segment (every CIGAR of the form [S]M[S]) — by far the common case.  All touch arithmetic
// collapses to adds and unsigned range checks in unit-position space; computed once per (tile, alignment) when the tile's chunk is
// staged in shared memory.
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_fast {
    u32 lo, span;        // touched positions [lo, lo + span]; for a simple alignment span == number of calls
    u32 lsrc_len;        // left-mate read offset of position lo | read length << 16
    u32 mlo, mlen;       // the right mate covers positions q with (q - mlo) < mlen ...
    u32 mdelta;          // ... where the mate position is q + mdelta            (all modulo 2^32, like the reference's unsigned math)
    u32 read;            // (read index << 1) | rc of the left mate
    u32 simple;          // 0: use ag_locate on the prepared record; bit 0: fast path valid; bit 1 (AG_FAST_CLEAN): every position the alignment
                         // touches — on the left mate and at the mate positions — holds at most one contiMer, so every touch has one candidate;
                         // bit 2 (AG_FAST_LINEAR): the contiMers under the mate positions it touches are one linear stretch of one contig thread
                         // (or there are none), so the mate-side match fields of every touch are (mcid0, q + mcd) without a table look-up
    u32 mcid0, mcd;      // see AG_FAST_LINEAR
    u32 aln, pad;        // alignment index (the prepared record of the generic path); 48 bytes = three 16-byte vectors
};
#define AG_FAST_CLEAN 2u
#define AG_FAST_LINEAR 4u

AG_HD ag_fast ag_fast_prep(const ag_alnp& p, u32 lo, u32 span, u32 aln_index = 0) {
    ag_fast f;
    f.mcid0 = AG_NONE; f.mcd = 0; f.aln = aln_index; f.pad = 0;
    f.lo = lo; f.span = span; f.read = p.left_read;
    u32 len = p.len_nseg & 0xFFFFu;
    f.simple = (((p.len_nseg >> 16) & 0xFF) == 1 && ((p.len_nseg >> 24) & 0xFF) == 1) ? 1u : 0u;
    u32 lsrc = p.l_sl & 0xFFFFu, rsrc = p.r_sl & 0xFFFFu;
    f.lsrc_len = lsrc | (len << 16);
    f.mlo = lo + rsrc - lsrc; f.mlen = p.r_sl >> 16; f.mdelta = p.r_dst - f.mlo;
    return f;
}

// many_prefix[p] = number of positions < p holding several contiMers (exclusive scan, n_pos + 1 entries)
AG_HD bool ag_fast_is_clean(const ag_fast& f, const ag_alnp& p, const u32* many_prefix) {
    if (!f.simple) return false;
    if (many_prefix[f.lo + f.span + 1] != many_prefix[f.lo]) return false;
    // right mate: the touch at q looks at r_dst + (q - mlo) for q in [lo, lo + span] with 0 <= q - mlo < mlen
    const long long mlo = (long long)f.lo + (long long)(p.r_sl & 0xFFFFu) - (long long)(p.l_sl & 0xFFFFu);
    const long long qa = (long long)f.lo > mlo ? (long long)f.lo : mlo;
    const long long qe = (long long)f.lo + f.span, me = mlo + (long long)f.mlen - 1;
    const long long qb = qe < me ? qe : me;
    if (qa > qb) return true;
    const u32 a = p.r_dst + (u32)(qa - mlo), b = p.r_dst + (u32)(qb - mlo);
    return many_prefix[b + 1] == many_prefix[a];
}

// lin_prefix[p] = number of positions 1 .. p-1 ... precisely: exclusive scan (n_pos + 1 entries) of brk[], where brk[p] = 1 when the contiMer
// summary of p does NOT continue that of p - 1 (both empty, or the same contig one offset further); brk[0] = 0
AG_HD u32 ag_cm1_break(const struct ag_cm1& prev, const struct ag_cm1& cur);
// clean + linear classification of a prepared alignment (k_prep): sets AG_FAST_CLEAN / AG_FAST_LINEAR and the linear mate fields
AG_HD void ag_fast_classify(ag_fast& f, const ag_alnp& p, const u32* many_prefix, const u32* lin_prefix, const struct ag_cm1* cm1);

// touch of a simple alignment at position q; requires q - f.lo <= f.span
AG_HD ag_touch ag_fast_touch(const ag_fast& f, u32 q, u32 k) {
    ag_touch t;
    u32 d = q - f.lo, len = f.lsrc_len >> 16, a = (f.lsrc_len & 0xFFFFu) + d;
    t.kind = d < f.span ? 1 : 2;
    t.soff = a; t.slen = t.kind == 1 ? k : ag_min_u32(k, len - a);
    t.mate = (q - f.mlo < f.mlen) ? q + f.mdelta : AG_NONE;
    t.npos = q + 1; t.nsoff = a + 1; t.nslen = ag_min_u32(k, len - (a + 1));
    t.nmate = (q + 1 - f.mlo < f.mlen) ? q + 1 + f.mdelta : AG_NONE;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------------
// reads: 2 bits per base (A0 C1 G2 T3), 16 bases per u32, fixed stride per read; 1-bit plane marks non-ACGT characters
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_reads {
    const u32* bases; const u32* nmask; const uint16_t* len;  // len per PAIR (mates are truncated to equal length, AG:3454)
    u32 stride2, stridem;
    // base code of the read in alignment orientation: 0-3, or 4 for anything the reference counts as 'N' (AG:1349)
    AG_HD int code(u32 read_rc, u32 rlen, u32 off) const {
        u32 read = read_rc >> 1, rc = read_rc & 1;
        u32 i = rc ? rlen - 1 - off : off;
        if ((nmask[(u64)read * stridem + (i >> 5)] >> (i & 31)) & 1) return 4;
        u32 c = (bases[(u64)read * stride2 + (i >> 4)] >> ((i & 15) * 2)) & 3;
        return rc ? 3 - (int)c : (int)c;
    }
};

// Oriented 4-bit codes of a read, eight per word: nibble t of word j = code (0-3, 4 = not ACGT) of ORIENTED offset 8j + t, i.e. what
// ag_reads::code(read_rc, len, 8j + t) returns; nibbles of offsets >= len are unspecified.  `b` / `m` = the read's packed words (2 bits per
// base / 1 bit per base), nb2 / nbm = their number.  Staging reads in this form turns the base lookup of a touch into one shift and mask.
AG_HD u32 ag_code4_word_ref(const u32* b, const u32* m, u32 rc, u32 len, u32 j) {   // plain per-base definition
    u32 w = 0;
    for (u32 t = 0; t < 8; t++) {
        const u32 soff = 8 * j + t;
        if (soff >= len) break;
        const u32 i = rc ? len - 1 - soff : soff;
        const u32 nb = (m[i >> 5] >> (i & 31)) & 1, c = (b[i >> 4] >> ((i & 15) * 2)) & 3;
        w |= (nb ? 4u : (rc ? 3u - c : c)) << (4 * t);
    }
    return w;
}
AG_HD u32 ag_brev32(u32 x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1); x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4); x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
AG_HD u32 ag_code4_word(const u32* b, const u32* m, u32 nb2, u32 nbm, u32 rc, u32 len, u32 j) {   // the same, eight bases at a time
    u32 x16, m8;
    if (!rc) {
        if (8 * j >= len) return 0;
        x16 = (b[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
        m8 = (m[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
    } else {
        if (8 * j + 7 >= len) return ag_code4_word_ref(b, m, rc, len, j);   // the read's last word on the reverse strand: fewer than eight bases
        const u32 lo = len - 8 - 8 * j;                                      // raw offsets lo .. lo + 7, to be reversed and complemented
        const u32 wi = lo >> 4, mi = lo >> 5;
        const u64 vb = (u64)b[wi] | (wi + 1 < nb2 ? (u64)b[wi + 1] << 32 : 0);
        const u64 vm = (u64)m[mi] | (mi + 1 < nbm ? (u64)m[mi + 1] << 32 : 0);
        u32 r = ag_brev32((u32)((vb >> ((lo & 15) * 2)) & 0xFFFFu)) >> 16;    // bit reversal reverses the groups but also swaps the two bits of each
        r = ((r & 0xAAAAu) >> 1) | ((r & 0x5555u) << 1);
        x16 = r ^ 0xFFFFu;                                                    // complement: 3 - c
        m8 = ag_brev32((u32)((vm >> (lo & 31)) & 0xFFu)) >> 24;
    }
    u32 x = x16; x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u;   // 2-bit groups -> nibbles
    u32 k = m8;  k = (k | (k << 12)) & 0x000F000Fu; k = (k | (k << 6)) & 0x03030303u; k = (k | (k << 3)) & 0x11111111u;   // mask bits -> bit 0 of the nibbles
    return (x & ~(k * 3u)) | (k << 2);
}

// ---------------------------------------------------------------------------------------------------------------------------
// candidates and the compatibility predicate
// ---------------------------------------------------------------------------------------------------------------------------
AG_HD bool ag_compatible(const ag_nodem& x, const ag_nodem& y, int iv) {  // AG:1293-1312 with OPTIMIZATION defined
    bool c1 = x.cid == AG_NONE || y.cid == AG_NONE || x.cid != y.cid || ag_absdiff(x.coff, y.coff) <= 5 * AG_EP;
    bool c2 = x.cid0 == AG_NONE || y.cid0 == AG_NONE || x.cid0 != y.cid0 || ag_absdiff(x.coff0, y.coff0) <= 2 * iv + 5 * AG_EP;
    bool c3 = x.moff == AG_NONE || y.moff == AG_NONE || ag_absdiff(x.moff, y.moff) <= 2 * iv + 5 * AG_EP;
    return c1 && c2 && c3;
}

// contiMers by position (static during the read phase)
struct ag_cmtab {
    const u32* start;  // CSR, n_pos + 1
    const ag_cm* cm;
    AG_HD u32 count(u32 pos) const { return start[pos + 1] - start[pos]; }
};

// cm1[pos] = (cid, coff) of the only contiMer at pos; (NONE, NONE) when there is none; (AG_CM_MANY, count) when there are several
// (then the CSR has to be walked).  One 8-byte load per mate lookup instead of three dependent ones.
#define AG_CM_MANY 0xFFFFFFFEu
struct ag_cm1 { u32 cid, coff; };
AG_HD ag_cm1 ag_make_cm1(const ag_cmtab& t, u32 pos) {
    u32 a = t.start[pos], n = t.start[pos + 1] - a;
    ag_cm1 r;
    if (n == 0) { r.cid = r.coff = AG_NONE; } else if (n == 1) { r.cid = t.cm[a].cid; r.coff = t.cm[a].coff; } else { r.cid = AG_CM_MANY; r.coff = n; }
    return r;
}

AG_HD u32 ag_cm1_break(const ag_cm1& prev, const ag_cm1& cur) {
    if (prev.cid == AG_NONE && cur.cid == AG_NONE) return 0u;
    if (prev.cid == AG_NONE || cur.cid == AG_NONE || prev.cid == AG_CM_MANY || cur.cid == AG_CM_MANY) return 1u;
    return (prev.cid == cur.cid && cur.coff == prev.coff + 1) ? 0u : 1u;
}
AG_HD void ag_fast_classify(ag_fast& f, const ag_alnp& p, const u32* many_prefix, const u32* lin_prefix, const ag_cm1* cm1) {
    if (!ag_fast_is_clean(f, p, many_prefix)) return;
    f.simple |= AG_FAST_CLEAN;
    // mate positions the touches look at (as in ag_fast_is_clean): r_dst + (q - mlo) for q in [lo, lo + span] with 0 <= q - mlo < mlen
    const long long mlo = (long long)f.lo + (long long)(p.r_sl & 0xFFFFu) - (long long)(p.l_sl & 0xFFFFu);
    const long long qa = (long long)f.lo > mlo ? (long long)f.lo : mlo;
    const long long qe = (long long)f.lo + f.span, me = mlo + (long long)f.mlen - 1;
    const long long qb = qe < me ? qe : me;
    if (qa > qb) { f.simple |= AG_FAST_LINEAR; f.mcid0 = AG_NONE; f.mcd = 0; return; }   // no touch has a mate position
    const u32 a = p.r_dst + (u32)(qa - mlo), b = p.r_dst + (u32)(qb - mlo);
    if (lin_prefix[b + 1] != lin_prefix[a + 1]) return;                                   // a break inside (a, b]
    const ag_cm1 c = cm1[a];
    f.simple |= AG_FAST_LINEAR; f.mcid0 = c.cid;
    f.mcd = c.cid == AG_NONE ? 0u : f.mdelta + c.coff - a;                                // coff0 of the touch at q = c.coff + (q + mdelta - a)
}

// Enumerate the candidates of a touch at `pos` with mate position `mate` in the reference's order: contiMers at pos (outer) x
// contiMers at the mate position (inner); an empty side contributes one "-1" entry (AG:1369-1477).
template <class F> AG_HD void ag_for_candidates(const ag_cmtab& t, u32 pos, u32 mate, F f) {
    u32 a0 = t.start[pos], na = t.start[pos + 1] - a0;
    u32 b0 = 0, nb = 0;
    if (mate != AG_NONE) { b0 = t.start[mate]; nb = t.start[mate + 1] - b0; }
    ag_nodem c; c.moff = mate;
    for (u32 ia = 0; ia < (na ? na : 1u); ia++) {
        if (na) { c.cid = t.cm[a0 + ia].cid; c.coff = t.cm[a0 + ia].coff; } else { c.cid = c.coff = AG_NONE; }
        for (u32 ib = 0; ib < (nb ? nb : 1u); ib++) {
            if (nb) { c.cid0 = t.cm[b0 + ib].cid; c.coff0 = t.cm[b0 + ib].coff; } else { c.cid0 = c.coff0 = AG_NONE; }
            f(c);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// node list of one position while it is being built
// ---------------------------------------------------------------------------------------------------------------------------
// The first AG_NODE_SCAP nodes of a position live in "slots" (shared memory on the device, [field][slot][thread], conflict-free);
// further nodes — and ALL nodes of the rare positions that hold several contiMers — are full ag_nodeb records chained through a
// global overflow pool.  A slot does not store contigID / contigOffset: on a position with at most one contiMer every candidate
// carries the same pair (the position's), so clause 1 of compatible() (AG:1296-1299) is identically true there.
#ifndef AG_NODE_SCAP
#define AG_NODE_SCAP 2
#endif
enum { AG_F_CID0 = 0, AG_F_COFF0 = 1, AG_F_MOFF = 2, AG_F_COV = 3, AG_F_CNT = 4, AG_F_SREAD = 9, AG_F_SL = 10, AG_F_SUCC = 11, AG_NF = 12 };
// A thread's slots.  On the device they sit in shared memory and are addressed through a 32-bit shared-space address with ld/st.shared
// (the address stays in one register; generic pointers into dynamic shared memory make the compiler re-derive the window base over and
// over); the host emulation uses a plain array.  Word index of (field, slot) = field * fstride + slot * nstride.
struct ag_slots {
#ifdef __CUDACC__   // nvcc (both passes see the same layout; the host pass never executes these)
    u32 saddr;
    static constexpr u32 fstride = AG_NODE_SCAP * AG_TILE, nstride = AG_TILE;
    __host__ __device__ __forceinline__ u32 ld(u32 field, u32 i) const {
        u32 v = 0;
#ifdef __CUDA_ARCH__
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + (field * fstride + i * nstride) * 4u));
#endif
        return v;
    }
    __host__ __device__ __forceinline__ void st(u32 field, u32 i, u32 v) const {
#ifdef __CUDA_ARCH__
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr + (field * fstride + i * nstride) * 4u), "r"(v) : "memory");
#endif
    }
#else
    u32* base; u32 fstride, nstride;
    inline u32 ld(u32 field, u32 i) const { return base[field * fstride + i * nstride]; }
    inline void st(u32 field, u32 i, u32 v) const { base[field * fstride + i * nstride] = v; }
#endif
    AG_HD void inc(u32 field, u32 i) const { st(field, i, ld(field, i) + 1u); }
};
struct ag_plist { u32 n, ovf_head, ovf_tail; };  // n counts slot nodes + pool nodes
struct ag_ovfpool { ag_nodeb* node; u32* next; u32* count; u32 cap; int* err; };

AG_HD bool ag_compat_b(const ag_nodem& c, const ag_nodeb& y, int iv) {
    ag_nodem m; m.cid = y.cid; m.coff = y.coff; m.cid0 = y.cid0; m.coff0 = y.coff0; m.moff = y.moff;
    return ag_compatible(c, m, iv);
}
// clauses 2 and 3 of compatible() (AG:1300-1310) on scalars
AG_HD bool ag_compat23(u32 cid0, u32 coff0, u32 moff, u32 ycid0, u32 ycoff0, u32 ymoff, int iv) {
    bool c2 = cid0 == AG_NONE || ycid0 == AG_NONE || cid0 != ycid0 || ag_absdiff(coff0, ycoff0) <= 2 * iv + 5 * AG_EP;
    bool c3 = moff == AG_NONE || ymoff == AG_NONE || ag_absdiff(moff, ymoff) <= 2 * iv + 5 * AG_EP;
    return c2 && c3;
}

// append a full record to the position's pool chain; returns the pool slot (NONE when the pool is exhausted: err is set and the
// whole sweep is repeated with a larger pool)
AG_HD u32 ag_pool_new(ag_plist& pl, const ag_ovfpool& pool, const ag_nodem& c, bool bump, int code, u32 sread, u32 soff_len) {
    u32 o = AG_ATOMIC_ADD(pool.count, 1u);
    if (o >= pool.cap) { AG_ATOMIC_OR(pool.err, 1); return AG_NONE; }   // bit E_OVF of the step's error word
    pool.next[o] = AG_NONE;
    if (pl.ovf_tail == AG_NONE) pl.ovf_head = o; else pool.next[pl.ovf_tail] = o;
    pl.ovf_tail = o;
    ag_nodeb* h = &pool.node[o];
    h->cid = c.cid; h->coff = c.coff; h->cid0 = c.cid0; h->coff0 = c.coff0; h->moff = c.moff;
    h->cov = bump ? 1u : 0u;
    for (u32 j = 0; j < 5; j++) h->cnt[j] = (bump && code == (int)j) ? 1u : 0u;
    h->sread = sread; h->soff_len = soff_len; h->succ = 0;
    return o;
}

// One candidate of one touch on a slot-mode position: first-compatible lookup, bump or create (AG:1375-1389 / AG:1493-1506).
// Returns the item index (stable: lists only grow at the tail).
AG_HD u32 ag_touch_slots(ag_plist& pl, const ag_slots& sv, const ag_ovfpool& pool, const ag_nodem& c, bool bump, int code, u32 sread, u32 soff_len, int iv) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (u32 i = 0; i < (u32)AG_NODE_SCAP; i++) {   // fixed trip count: slot addresses become immediates
        if (i >= pl.n) break;
        if (ag_compat23(c.cid0, c.coff0, c.moff, sv.ld(AG_F_CID0, i), sv.ld(AG_F_COFF0, i), sv.ld(AG_F_MOFF, i), iv)) {
            if (bump) { sv.inc(AG_F_COV, i); if (code >= 0) sv.inc(AG_F_CNT + (u32)code, i); }
            return i;
        }
    }
    if (pl.n > (u32)AG_NODE_SCAP) {
        u32 i = AG_NODE_SCAP;
        for (u32 o = pl.ovf_head; o != AG_NONE; o = pool.next[o], i++) {
            ag_nodeb* h = &pool.node[o];
            if (ag_compat23(c.cid0, c.coff0, c.moff, h->cid0, h->coff0, h->moff, iv)) { if (bump) { h->cov++; if (code >= 0) h->cnt[code]++; } return i; }
        }
    }
    const u32 i = pl.n;
    if (i < (u32)AG_NODE_SCAP) {
        sv.st(AG_F_CID0, i, c.cid0); sv.st(AG_F_COFF0, i, c.coff0); sv.st(AG_F_MOFF, i, c.moff);
        sv.st(AG_F_COV, i, bump ? 1u : 0u);
        for (u32 j = 0; j < 5; j++) sv.st(AG_F_CNT + j, i, (bump && code == (int)j) ? 1u : 0u);
        sv.st(AG_F_SREAD, i, sread); sv.st(AG_F_SL, i, soff_len); sv.st(AG_F_SUCC, i, 0u);
    } else if (ag_pool_new(pl, pool, c, bump, code, sread, soff_len) == AG_NONE) return i;
    pl.n = i + 1;
    return i;
}

// the same on a position that holds several contiMers: every node is a pool record and all three clauses are evaluated
AG_HD u32 ag_touch_pool(ag_plist& pl, const ag_ovfpool& pool, const ag_nodem& c, bool bump, int code, u32 sread, u32 soff_len, int iv) {
    u32 i = 0;
    for (u32 o = pl.ovf_head; o != AG_NONE; o = pool.next[o], i++)
        if (ag_compat_b(c, pool.node[o], iv)) { ag_nodeb* h = &pool.node[o]; if (bump) { h->cov++; if (code >= 0) h->cnt[code]++; } return i; }
    if (ag_pool_new(pl, pool, c, bump, code, sread, soff_len) != AG_NONE) pl.n = i + 1;
    return i;
}

// record "the call that bumped item `item` of this (slot-mode) position continues on item `nb` of the next position"
AG_HD void ag_note_succ(const ag_plist& pl, const ag_slots& sv, const ag_ovfpool& pool, u32 item, u32 nb) {
    if (item < (u32)AG_NODE_SCAP) { sv.st(AG_F_SUCC, item, sv.ld(AG_F_SUCC, item) | (1u << nb)); return; }
    u32 o = pl.ovf_head;                                   // rare: the item lives in the pool chain
    for (u32 i = AG_NODE_SCAP; i < item && o != AG_NONE; i++) o = pool.next[o];
    if (o != AG_NONE) pool.node[o].succ |= 1u << nb;        // (NONE: pool exhausted, the sweep is repeated)
}

// What one lane (= one unit position q) does with one alignment of its tile during the node sweep.  Result:
//   want:   a call starts at q (kind-1 touch), so an edge to the alignment's item at the call's successor position is due;
//   return: the item this lane resolved, offered to the lane on its left as that successor item — only for CLEAN alignments (successor
//           position = q + 1, i.e. the next lane, and exactly one candidate per touch); NONE otherwise.
// An edge that cannot be settled through (want, item of the right neighbour) flags the tile for the generic edge sweep: flag 1 when the
// alignment is not clean (the generic sweep then handles exactly the non-clean alignments of the tile), flag 2 when a clean alignment
// met an item index >= 32 (the generic sweep then redoes every call of the tile; edges are de-duplicated).

// rare cases (several contiMers on either side, multi-segment CIGARs, gap chains): the reference's nested candidate enumeration
// (AG:1369-1477).  Out of line and by value so that the hot path keeps its list state in registers.
struct ag_gen_ret { ag_plist pl; bool want; };
template <class CodeF>
AG_HD_COLD ag_gen_ret ag_lane_generic(ag_plist pl, const ag_slots sv, const ag_ovfpool pool, const ag_cmtab cmt, const ag_cm1 ca, const ag_fast f, const ag_alnp* ap,
                                      const ag_seg* ext, u32 q, u32 k, int iv, bool force_generic, CodeF codef) {
    ag_gen_ret r; r.pl = pl; r.want = false;
    ag_touch t;
    if (f.simple && !force_generic) t = ag_fast_touch(f, q, k);
    else { t = ag_locate(*ap, ext, q, k); if (!t.kind) return r; }
    int code = -1;
    if (t.kind == 1 && t.slen) code = codef(t.soff);
    const u32 sl = t.soff | (t.slen << 16);
    const bool bump = t.kind == 1;
    const bool slots = ca.cid != AG_CM_MANY;
    ag_for_candidates(cmt, q, t.mate, [&](const ag_nodem& c) {
        if (slots) ag_touch_slots(pl, sv, pool, c, bump, code, f.read, sl, iv);
        else ag_touch_pool(pl, pool, c, bump, code, f.read, sl, iv);
    });
    r.pl = pl; r.want = bump;
    return r;
}

AG_HD u32 ag_fast_mate(const ag_fast& f, u32 q) { return (q - f.mlo < f.mlen) ? q + f.mdelta : AG_NONE; }

// everything one lane does with one tile alignment whose touch range contains q (q - f.lo <= f.span)
template <class CodeF>
AG_HD u32 ag_lane_touch(bool& want, ag_plist& pl, const ag_slots& sv, const ag_ovfpool& pool, const ag_cmtab& cmt, const ag_cm1* cm1, const ag_cm1& ca, const ag_fast& f,
                        const ag_alnp* ap, const ag_seg* ext, u32 q, u32 k, int iv, bool force_generic, CodeF codef) {
    if ((f.simple & AG_FAST_CLEAN) && !force_generic) {
        // the common case: one M segment per mate and at most one contiMer at q and at the mate position => exactly one candidate
        const u32 mate = ag_fast_mate(f, q);
        ag_nodem c; c.cid = ca.cid; c.coff = ca.coff; c.cid0 = c.coff0 = AG_NONE; c.moff = mate;
        if (mate != AG_NONE) {
            if (f.simple & AG_FAST_LINEAR) { c.cid0 = f.mcid0; c.coff0 = f.mcid0 != AG_NONE ? q + f.mcd : AG_NONE; }
            else { const ag_cm1 cb = cm1[mate]; c.cid0 = cb.cid; c.coff0 = cb.coff; }
        }
        const u32 d = q - f.lo, len = f.lsrc_len >> 16, a = (f.lsrc_len & 0xFFFFu) + d;
        const bool bump = d < f.span;                             // a call starts here (kind 1); else the stand-alone k2 of the last call
        const u32 slen = bump ? k : ag_min_u32(k, len - a);
        int code = -1;
        if (bump && slen) code = codef(a);
        want = bump;
        return ag_touch_slots(pl, sv, pool, c, bump, code, f.read, a | (slen << 16), iv);
    }
    const ag_gen_ret r = ag_lane_generic(pl, sv, pool, cmt, ca, f, ap, ext, q, k, iv, force_generic, codef);
    pl = r.pl; want = r.want;
    return AG_NONE;
}

// first node of the FINAL list [nb, nb + n) compatible with candidate c
AG_HD u32 ag_first_compatible(const ag_nodec* nc, const ag_nodew* nw, u32 nb, u32 n, const ag_nodem& c, int iv) {
    for (u32 i = 0; i < n; i++) {
        const ag_nodec y = nc[nb + i];
        ag_nodem m; m.cid = y.cid; m.coff = y.coff; m.cid0 = y.cid0; m.coff0 = y.coff0; m.moff = nw[nb + i].moff;
        if (ag_compatible(c, m, iv)) return i;
    }
    return AG_NONE;
}
AG_HD bool ag_edge_ok_c(const ag_nodec& x, const ag_nodec& y, int iv) {  // AG:1600-1615
    bool c1 = y.cid == AG_NONE || x.cid == AG_NONE || y.cid != x.cid || ag_absdiff(y.coff, x.coff) <= 5 * AG_EP;
    bool c2 = y.cid0 == AG_NONE || x.cid0 == AG_NONE || y.cid0 != x.cid0 || ag_absdiff(y.coff0, x.coff0) <= 2 * iv + 5 * AG_EP;
    return c1 && c2;
}

// consensus base of a node (AG:1944-1952 + AG:1997-2001)
AG_HD char ag_consensus(const u32* cnt, char refbase) {
    u32 a = cnt[0], c = cnt[1], g = cnt[2], t = cnt[3], n = cnt[4];
    if (!a && !c && !g && !t && !n) return refbase;
    if (a >= c && a >= g && a >= t && a >= n) return 'A';
    if (c >= a && c >= g && c >= t && c >= n) return 'C';
    if (g >= a && g >= c && g >= t && g >= n) return 'G';
    if (t >= a && t >= c && t >= g && t >= n) return 'T';
    return 'N';
}

// misc word of a final node: consensus base, coverage filter (AG:1912-1915), contigOffset != -1 (AG:2004)
AG_HD u32 ag_node_misc(u32 cid, u32 coff, u32 cov, const u32* cnt, char refbase, int coverage) {
    u32 misc = (u32)(unsigned char)ag_consensus(cnt, refbase);
    if (cid == AG_NONE && (int)cov < coverage) misc |= AG_NW_FILTERED | AG_NW_TRAV;
    if (coff != AG_NONE) misc |= AG_NW_HASCONTIG;
    return misc;
}

// ---------------------------------------------------------------------------------------------------------------------------
// extension walk (AG:1954-2204) on the final table
// ---------------------------------------------------------------------------------------------------------------------------
struct ag_walkctx {
    ag_nodew* nw;             // per node; misc carries the traversed / detour marks
    const u32* node_pos;      // node -> unit position
    const u32* pos_node;      // CSR position -> first node index (n_pos + 1)
    const u32* ovf_head;      // per node: head of overflow successor list (valid when misc & AG_NW_OVF)
    const u32* ovf_target;    // overflow pool
    const u32* ovf_next;
    ag_cmtab cmt;
    const u32* chain_pos;     // chain-major: unit position of every contiMer
    u32* walk_next;           // per node: next node of the walk that marked it, or NONE
    const ag_chain* chain;    // forced-link chains
    const ag_hrec* hrec;      // per chain head: packed hop record (valid for live, non-interior nodes)
    const ag_hdet* hdet;      // per chain head with hrec.tcm != NONE: where the contiMer detour of its tail lands (precomputed: the replay's longest dependent chain)
    // exact sequential replay (skip rule, AG:2194-2202): a chain can be entered at an interior node, so marks are kept as a marked SUFFIX
    // per chain, indexed by the chain's tail: msuf = nodes marked at the tail end, mnode = first marked node
    u32* msuf; u32* mnode; const u32* fprev;
};

// untraversed successors of node v (record `nd` already loaded): count, `pick` = the last one seen, `prec` = its record (AG:2020-2032)
AG_HD u32 ag_live_succ(const ag_walkctx& w, u32 v, const ag_nodew& nd, u32& pick, ag_nodew& prec) {
    u32 cnt = 0; pick = AG_NONE;
    ag_nodew s0, s1;
    if (nd.succ0 != AG_NONE) s0 = w.nw[nd.succ0];
    if (nd.succ1 != AG_NONE) s1 = w.nw[nd.succ1];
    if (nd.succ0 != AG_NONE && !(s0.misc & AG_NW_TRAV)) { cnt++; pick = nd.succ0; prec = s0; }
    if (nd.succ1 != AG_NONE && !(s1.misc & AG_NW_TRAV)) { cnt++; pick = nd.succ1; prec = s1; }
    if (nd.misc & AG_NW_OVF)
        for (u32 o = w.ovf_head[v]; o != AG_NONE; o = w.ovf_next[o]) {
            u32 s = w.ovf_target[o]; ag_nodew r = w.nw[s];
            if (!(r.misc & AG_NW_TRAV)) { cnt++; pick = s; prec = r; }
        }
    return cnt;
}

// Forced link u -> w: w is u's only live successor, u is w's only live predecessor, and neither sits on a position
// that holds a terminal contiMer (where a detour can inspect nodes and jump past them, AG:2093-2136).  Then w is marked exactly
// when u is, by the same walk, so chains of forced links can be contracted.
AG_HD u32 ag_forced_succ(const ag_nodew* nw, const u32* ovf_head, const u32* ovf_target, const u32* ovf_next, const u32* indeg, const unsigned char* pos_term,
                         const u32* node_pos, u32 u) {
    const ag_nodew nd = nw[u];
    if (nd.misc & AG_NW_FILTERED) return AG_NONE;
    if (pos_term[node_pos[u]]) return AG_NONE;
    u32 cnt = 0, w = AG_NONE;
    if (nd.succ0 != AG_NONE && !(nw[nd.succ0].misc & AG_NW_FILTERED)) { cnt++; w = nd.succ0; }
    if (nd.succ1 != AG_NONE && !(nw[nd.succ1].misc & AG_NW_FILTERED)) { cnt++; w = nd.succ1; }
    if (nd.misc & AG_NW_OVF) for (u32 o = ovf_head[u]; o != AG_NONE; o = ovf_next[o]) if (!(nw[ovf_target[o]].misc & AG_NW_FILTERED)) { cnt++; w = ovf_target[o]; }
    if (cnt != 1 || indeg[w] != 1 || pos_term[node_pos[w]]) return AG_NONE;
    return w;
}

// Simulate the walk that starts at untraversed chain head `start`.  Marks nodes, records the path in walk_next / the detour bit.
// A chain of forced links is marked as one (only its head carries the mark: interior nodes are never inspected by anyone else).
AG_HD ag_walk ag_walk_from(const ag_walkctx& w, u32 start) {
    ag_walk r;
    const ag_nodew sn = w.nw[start];
    r.start_node = start; r.soff = w.node_pos[start]; r.soff0 = sn.moff;
    ag_hrec h = w.hrec[start];
    u32 v = start, vmisc = sn.misc, len = 0, ext = 0, t;
    for (;;) {
        // kMerTag == 1 steps over the chain headed by v (AG:1997-2060)
        vmisc |= AG_NW_TRAV;
        w.nw[v].misc = vmisc;
        len += h.len;
        if (h.flg) ext = 1;
        t = h.tail;
        const u32 tmisc = (t == v) ? vmisc : h.tmisc;
        // untraversed successors of the tail (AG:2020-2032); the hop records of both candidates are fetched with their marks
        u32 cnt = 0, pick = AG_NONE, pmisc = 0;
        ag_hrec ph = h;
        {
            ag_nodew r0, r1; ag_hrec g0 = h, g1 = h;
            r0.misc = r1.misc = AG_NW_TRAV;
            if (h.ts0 != AG_NONE) { r0 = w.nw[h.ts0]; g0 = w.hrec[h.ts0]; }
            if (h.ts1 != AG_NONE) { r1 = w.nw[h.ts1]; g1 = w.hrec[h.ts1]; }
            if (!(r0.misc & AG_NW_TRAV)) { cnt++; pick = h.ts0; pmisc = r0.misc; ph = g0; }
            if (!(r1.misc & AG_NW_TRAV)) { cnt++; pick = h.ts1; pmisc = r1.misc; ph = g1; }
            if (tmisc & AG_NW_OVF)
                for (u32 o = w.ovf_head[t]; o != AG_NONE; o = w.ovf_next[o]) {
                    const u32 s = w.ovf_target[o]; const ag_nodew rr = w.nw[s];
                    if (!(rr.misc & AG_NW_TRAV)) { cnt++; pick = s; pmisc = rr.misc; ph = w.hrec[s]; }
                }
        }
        if (cnt == 1) { w.walk_next[t] = pick; v = pick; vmisc = pmisc; h = ph; continue; }
        if (h.tcm != AG_NONE) {
            // switch to the contiMer thread (AG:2047-2057), run to its terminal (AG:2064-2072), try to re-enter (AG:2093-2136)
            const ag_hdet dt = w.hdet[v];   // (v heads the chain whose tail is t)
            len += dt.len; ext = 1;
            w.nw[t].misc = tmisc | AG_NW_DETOUR;
            const u32 z = dt.z;
            u32 live = 0, item = AG_NONE; ag_nodew irec;
            for (u32 x = dt.first; x < dt.first + dt.n; x++) { ag_nodew xr = w.nw[x]; if (!(xr.misc & AG_NW_TRAV)) { live++; item = x; irec = xr; } }
            u32 pick2 = AG_NONE, cnt2 = 0; ag_nodew prec2;
            if (live == 1) cnt2 = ag_live_succ(w, item, irec, pick2, prec2);
            if (cnt2 == 1) { w.walk_next[t] = pick2; v = pick2; vmisc = prec2.misc; h = w.hrec[pick2]; continue; }
            w.walk_next[t] = AG_NONE;
            r.eoff = z; r.eoff0 = AG_NONE; r.flags = ext | (1u << 1);  // kMerTag -2
            break;
        }
        w.walk_next[t] = AG_NONE;
        r.eoff = w.node_pos[t]; r.eoff0 = h.tmoff; r.flags = ext | (0u << 1);  // kMerTag -1
        break;
    }
    r.len = len; r.last_node = t;
    return r;
}

// hop record of chain head v (k_hrec); p = unit position of the chain's tail
AG_HD ag_hrec ag_make_hrec(const ag_chain& c, const ag_nodew& tail, const ag_cmtab& cmt, u32 p) {
    ag_hrec h; h.tail = c.tail; h.len = c.len; h.flg = c.flg; h.ts0 = tail.succ0; h.ts1 = tail.succ1; h.tmisc = tail.misc; h.tmoff = tail.moff;
    const u32 c0 = cmt.start[p];
    h.tcm = (cmt.start[p + 1] - c0 == 1 && cmt.cm[c0].chain != cmt.cm[c0].term) ? c0 : AG_NONE;
    return h;
}

// detour record of a chain head whose hop record has tcm != NONE (k_hrec)
AG_HD ag_hdet ag_make_hdet(const ag_cmtab& cmt, const u32* chain_pos, const u32* pos_node, u32 tcm) {
    const ag_cm m = cmt.cm[tcm];
    ag_hdet dt; dt.z = chain_pos[m.term]; dt.first = pos_node[dt.z]; dt.n = pos_node[dt.z + 1] - dt.first; dt.len = m.term - m.chain;
    return dt;
}

// ---- exact sequential replay with chains ----------------------------------------------------------------------------------
// With the 1000-position skip (AG:2194-2202) the scan can start a walk INSIDE a chain (its head was skipped), so a chain is no longer
// all-or-nothing; but walks only run forward, so the marked nodes of a chain always form a suffix.  traversed(x) <=> x is filtered or
// lies in the marked suffix of its chain.
AG_HD bool ag_seq_trav(const ag_walkctx& w, u32 x) {
    if (w.nw[x].misc & AG_NW_FILTERED) return true;
    const ag_chain c = w.chain[x];
    return c.len <= w.msuf[c.tail];
}
AG_HD u32 ag_seq_live_succ(const ag_walkctx& w, u32 v, const ag_nodew& nd, u32& pick) {
    u32 cnt = 0; pick = AG_NONE;
    if (nd.succ0 != AG_NONE && !ag_seq_trav(w, nd.succ0)) { cnt++; pick = nd.succ0; }
    if (nd.succ1 != AG_NONE && !ag_seq_trav(w, nd.succ1)) { cnt++; pick = nd.succ1; }
    if (nd.misc & AG_NW_OVF) for (u32 o = w.ovf_head[v]; o != AG_NONE; o = w.ovf_next[o]) { u32 s = w.ovf_target[o]; if (!ag_seq_trav(w, s)) { cnt++; pick = s; } }
    return cnt;
}
AG_HD ag_walk ag_walk_from_seq(const ag_walkctx& w, u32 start) {
    ag_walk r;
    r.start_node = start; r.soff = w.node_pos[start]; r.soff0 = w.nw[start].moff;
    u32 v = start, len = 0, ext = 0, stop = start;
    for (;;) {
        const ag_chain c = w.chain[v];
        const u32 t = c.tail, ms = w.msuf[t], old_m = w.mnode[t];
        const u32 sufflg = ms ? w.chain[old_m].flg : 0;
        len += c.len - ms;
        if (c.flg - sufflg) ext = 1;
        w.msuf[t] = c.len; w.mnode[t] = v;                  // the chain is now marked from v to its tail
        stop = ms ? w.fprev[old_m] : t;                     // last node this walk marks in the chain
        const ag_nodew sn = w.nw[stop];
        u32 pick = AG_NONE, cnt = 0;
        if (!ms) cnt = ag_seq_live_succ(w, stop, sn, pick); // otherwise stop's only live successor is the (traversed) old suffix start
        if (cnt == 1) { w.nw[stop].misc = sn.misc | AG_NW_STOP; w.walk_next[stop] = pick; v = pick; continue; }
        const u32 p = w.node_pos[stop], c0 = w.cmt.start[p];
        if (w.cmt.start[p + 1] - c0 == 1 && w.cmt.cm[c0].chain != w.cmt.cm[c0].term) {
            const ag_cm m = w.cmt.cm[c0];
            len += m.term - m.chain; ext = 1;
            w.nw[stop].misc = sn.misc | AG_NW_STOP | AG_NW_DETOUR;
            const u32 z = w.chain_pos[m.term];
            u32 live = 0, item = AG_NONE;
            for (u32 x = w.pos_node[z]; x < w.pos_node[z + 1]; x++) if (!ag_seq_trav(w, x)) { live++; item = x; }
            u32 pick2 = AG_NONE, cnt2 = 0;
            if (live == 1) cnt2 = ag_seq_live_succ(w, item, w.nw[item], pick2);
            if (cnt2 == 1) { w.walk_next[stop] = pick2; v = pick2; continue; }
            w.walk_next[stop] = AG_NONE;
            r.eoff = z; r.eoff0 = AG_NONE; r.flags = ext | (1u << 1);
            break;
        }
        w.nw[stop].misc = sn.misc | AG_NW_STOP;
        w.walk_next[stop] = AG_NONE;
        r.eoff = p; r.eoff0 = sn.moff; r.flags = ext | (0u << 1);
        break;
    }
    r.len = len; r.last_node = stop;
    return r;
}

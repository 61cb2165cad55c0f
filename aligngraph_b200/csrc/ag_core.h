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
#else
template <class T> static inline T ag_host_add(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T ag_host_min(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T ag_host_max(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T ag_host_cas(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
#define AG_ATOMIC_ADD(p, v) ag_host_add((p), (v))
#define AG_ATOMIC_MIN(p, v) ag_host_min((p), (v))
#define AG_ATOMIC_MAX(p, v) ag_host_max((p), (v))
#define AG_ATOMIC_CAS(p, c, v) ag_host_cas((p), (c), (v))
#endif

#define AG_TILE 256  // unit positions per tile (= threads per CTA of the node / edge sweeps)

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

AG_HD u32 ag_min_u32(u32 a, u32 b) { return a < b ? a : b; }

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
    u32 simple;          // 1: fast path valid; 0: use ag_locate on the prepared record
};

AG_HD ag_fast ag_fast_prep(const ag_alnp& p, u32 lo, u32 span) {
    ag_fast f;
    f.lo = lo; f.span = span; f.read = p.left_read;
    u32 len = p.len_nseg & 0xFFFFu;
    f.simple = (((p.len_nseg >> 16) & 0xFF) == 1 && ((p.len_nseg >> 24) & 0xFF) == 1) ? 1u : 0u;
    u32 lsrc = p.l_sl & 0xFFFFu, rsrc = p.r_sl & 0xFFFFu;
    f.lsrc_len = lsrc | (len << 16);
    f.mlo = lo + rsrc - lsrc; f.mlen = p.r_sl >> 16; f.mdelta = p.r_dst - f.mlo;
    return f;
}

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

// ---------------------------------------------------------------------------------------------------------------------------
// candidates and the compatibility predicate
// ---------------------------------------------------------------------------------------------------------------------------
AG_HD bool ag_compatible(const ag_nodem& x, const ag_nodem& y, int iv) {  // AG:1293-1312 with OPTIMIZATION defined
    bool c1 = x.cid == AG_NONE || y.cid == AG_NONE || x.cid != y.cid || ag_absdiff(x.coff, y.coff) <= 5 * AG_EP;
    bool c2 = x.cid0 == AG_NONE || y.cid0 == AG_NONE || x.cid0 != y.cid0 || ag_absdiff(x.coff0, y.coff0) <= 2 * iv + 5 * AG_EP;
    bool c3 = x.moff == AG_NONE || y.moff == AG_NONE || ag_absdiff(x.moff, y.moff) <= 2 * iv + 5 * AG_EP;
    return c1 && c2 && c3;
}
AG_HD bool ag_edge_ok(const ag_nodem& x, const ag_nodem& y, int iv) {  // AG:1600-1615
    bool c1 = y.cid == AG_NONE || x.cid == AG_NONE || y.cid != x.cid || ag_absdiff(y.coff, x.coff) <= 5 * AG_EP;
    bool c2 = y.cid0 == AG_NONE || x.cid0 == AG_NONE || y.cid0 != x.cid0 || ag_absdiff(y.coff0, x.coff0) <= 2 * iv + 5 * AG_EP;
    return c1 && c2;
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
#define AG_NODE_CAP 6
struct ag_ovfpool { ag_nodeb* node; u32* next; u32* count; u32 cap; int* err; };

AG_HD bool ag_compat_b(const ag_nodem& c, const ag_nodeb& y, int iv) {
    ag_nodem m; m.cid = y.cid; m.coff = y.coff; m.cid0 = y.cid0; m.coff0 = y.coff0; m.moff = y.moff;
    return ag_compatible(c, m, iv);
}

// Strided view of a position's node list: field f of node i lives at base[f * fstride + i * nstride].  On the device the first
// `cap` nodes of every position sit in shared memory as [field][node][thread] (conflict-free); the host emulation / generic code
// uses the same accessors over an array of ag_nodeb (fstride 1, nstride 13).  Nodes beyond `cap` go to a global overflow pool.
struct ag_nview {
    u32* base; u32 fstride, nstride, cap;
    u32 n, ovf_head, ovf_tail;
    AG_HD void init(u32* b, u32 fs, u32 ns, u32 c) { base = b; fstride = fs; nstride = ns; cap = c; n = 0; ovf_head = ovf_tail = AG_NONE; }
    AG_HD u32& f(u32 field, u32 i) const { return base[field * fstride + i * nstride]; }
    AG_HD ag_nodem match(u32 i) const { ag_nodem m; m.cid = f(0, i); m.coff = f(1, i); m.cid0 = f(2, i); m.coff0 = f(3, i); m.moff = f(4, i); return m; }
    AG_HD ag_nodeb get(u32 i) const {
        ag_nodeb b; b.cid = f(0, i); b.coff = f(1, i); b.cid0 = f(2, i); b.coff0 = f(3, i); b.moff = f(4, i); b.cov = f(5, i);
        for (u32 j = 0; j < 5; j++) b.cnt[j] = f(6 + j, i);
        b.sread = f(11, i); b.soff_len = f(12, i);
        return b;
    }
};

// One candidate of one touch: first-compatible lookup, bump or create  (AG:1375-1389 / AG:1493-1506)
AG_HD void ag_node_touch_v(ag_nview& nl, const ag_ovfpool& pool, const ag_nodem& c, bool bump, int code, u32 sread, u32 soff_len, int iv) {
    u32 nloc = nl.n < nl.cap ? nl.n : nl.cap;
    for (u32 i = 0; i < nloc; i++)
        if (ag_compatible(c, nl.match(i), iv)) { if (bump) { nl.f(5, i)++; if (code >= 0) nl.f(6 + (u32)code, i)++; } return; }
    if (nl.n > nl.cap)
        for (u32 o = nl.ovf_head; o != AG_NONE; o = pool.next[o])
            if (ag_compat_b(c, pool.node[o], iv)) { if (bump) { pool.node[o].cov++; if (code >= 0) pool.node[o].cnt[code]++; } return; }
    if (nl.n < nl.cap) {
        u32 i = nl.n;
        nl.f(0, i) = c.cid; nl.f(1, i) = c.coff; nl.f(2, i) = c.cid0; nl.f(3, i) = c.coff0; nl.f(4, i) = c.moff;
        nl.f(5, i) = bump ? 1u : 0u;
        for (u32 j = 0; j < 5; j++) nl.f(6 + j, i) = (bump && code == (int)j) ? 1u : 0u;
        nl.f(11, i) = sread; nl.f(12, i) = soff_len;
    } else {
        u32 o = AG_ATOMIC_ADD(pool.count, 1u);
        if (o >= pool.cap) { *pool.err = 1; return; }
        pool.next[o] = AG_NONE;
        if (nl.ovf_tail == AG_NONE) nl.ovf_head = o; else pool.next[nl.ovf_tail] = o;
        nl.ovf_tail = o;
        ag_nodeb* h = &pool.node[o];
        h->cid = c.cid; h->coff = c.coff; h->cid0 = c.cid0; h->coff0 = c.coff0; h->moff = c.moff;
        h->cov = bump ? 1u : 0u;
        for (u32 j = 0; j < 5; j++) h->cnt[j] = (bump && code == (int)j) ? 1u : 0u;
        h->sread = sread; h->soff_len = soff_len;
    }
    nl.n++;
}

// first node of a FINAL list compatible with candidate c
AG_HD u32 ag_first_compatible(const ag_nodem* nodes, u32 n, const ag_nodem& c, int iv) {
    for (u32 i = 0; i < n; i++) if (ag_compatible(c, nodes[i], iv)) return i;
    return AG_NONE;
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

// Simulate the walk that starts at untraversed node `start`.  Marks nodes, records the path in walk_next / the detour bit.
AG_HD ag_walk ag_walk_from(const ag_walkctx& w, u32 start) {
    ag_walk r;
    ag_nodew cur = w.nw[start];
    r.start_node = start; r.soff = w.node_pos[start]; r.soff0 = cur.moff;
    u32 v = start, len = 0, ext = 0;
    for (;;) {
        // kMerTag == 1 step on an untraversed node (AG:1997-2060)
        len++;
        if (cur.misc & AG_NW_HASCONTIG) ext = 1;
        cur.misc |= AG_NW_TRAV;
        w.nw[v].misc = cur.misc;
        if (w.chain) {  // a chain of forced links is marked as one: jump to its tail (interior nodes are never inspected by anyone else)
            ag_chain c = w.chain[v];
            if (c.tail != v) { len += c.len - 1; if (c.flg) ext = 1; v = c.tail; cur = w.nw[v]; }
        }
        u32 pick; ag_nodew prec;
        u32 cnt = ag_live_succ(w, v, cur, pick, prec);
        if (cnt == 1) { w.walk_next[v] = pick; v = pick; cur = prec; continue; }
        u32 p = w.node_pos[v];
        u32 c0 = w.cmt.start[p];
        if (w.cmt.start[p + 1] - c0 == 1 && w.cmt.cm[c0].chain != w.cmt.cm[c0].term) {
            // switch to the contiMer thread (AG:2047-2057), run to its terminal (AG:2064-2072), try to re-enter (AG:2093-2136)
            ag_cm m = w.cmt.cm[c0];
            len += m.term - m.chain; ext = 1;
            w.nw[v].misc = cur.misc | AG_NW_DETOUR;
            u32 z = w.chain_pos[m.term];
            u32 live = 0, item = AG_NONE; ag_nodew irec;
            for (u32 x = w.pos_node[z]; x < w.pos_node[z + 1]; x++) { ag_nodew xr = w.nw[x]; if (!(xr.misc & AG_NW_TRAV)) { live++; item = x; irec = xr; } }
            u32 pick2 = AG_NONE, cnt2 = 0; ag_nodew prec2;
            if (live == 1) cnt2 = ag_live_succ(w, item, irec, pick2, prec2);
            if (cnt2 == 1) { w.walk_next[v] = pick2; v = pick2; cur = prec2; continue; }
            w.walk_next[v] = AG_NONE;
            r.eoff = z; r.eoff0 = AG_NONE; r.flags = ext | (1u << 1);  // kMerTag -2
            break;
        }
        w.walk_next[v] = AG_NONE;
        r.eoff = p; r.eoff0 = cur.moff; r.flags = ext | (0u << 1);  // kMerTag -1
        break;
    }
    r.len = len; r.last_node = v;
    return r;
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

// SAM record -> alignment segments, as pure functions over bytes (no allocation, no library calls) that compile for the host and for the
// device.  The host parsers (ag_host.cpp) use them today; they are written this way so that the same code can run one thread per record
// pair on the GPU.  Semantics: parseBOWTIE, AlignGraph.cpp:181-285 (cited AG:line).
#pragma once
#include "ag_types.h"

#define AG_SAM_MAXSEG 24   // M segments kept per record; a CIGAR with more sets AG_SAM_ERR_SEGS and the caller falls back to the general path

enum { AG_SAM_OK = 0, AG_SAM_ERR_CHAR = 1, AG_SAM_ERR_SEGS = 2 };

struct ag_samline {
    u32 sid, fr;                                     // QNAME as an integer (AG:197), FLAG & 0x10 (AG:199)
    u32 tid, tstart, tend, tgap;                     // RNAME id (NONE = unaligned, AG:203-209), target interval, deleted bases
    u32 sstart, send, sgap, ssize;                   // source interval (soft clips excluded), inserted bases, read length according to the CIGAR
    u32 nseg; ag_seg seg[AG_SAM_MAXSEG];             // M segments in CIGAR order: read offset, unit position (0-based), length
    u32 err; char bad_char;                          // AG_SAM_ERR_CHAR: `unknown character: X` (AG:263-269)
};

// atoi() over a bounded field (leading white space, optional sign, digits; saturates like strtol) — the reference calls atoi on every field
AG_HD int ag_sam_atoi(const char* s, u32 n) {
    u32 i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    unsigned long long v = 0; bool sat = false;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) { v = v * 10 + (unsigned)(s[i] - '0'); if (v > 0x7FFFFFFFFFFFFFFFull) sat = true; }
    const long long r = sat ? (neg ? (long long)0x8000000000000000ull : 0x7FFFFFFFFFFFFFFFll) : (neg ? -(long long)v : (long long)v);
    return (int)r;
}

AG_HD void ag_sam_parse_line(const char* s, u32 n, ag_samline& r) {
    // the first six tab-separated fields: QNAME FLAG RNAME POS MAPQ CIGAR
    u32 fo[6], fl[6]; u32 nf = 0, p = 0;
    while (nf < 6) {
        u32 e = p;
        while (e < n && s[e] != '\t') e++;
        fo[nf] = p; fl[nf] = e - p; nf++;
        if (e >= n) break;
        p = e + 1;
    }
    for (u32 i = nf; i < 6; i++) { fo[i] = n; fl[i] = 0; }
    r.err = AG_SAM_OK; r.bad_char = 0; r.nseg = 0;
    r.sid = (u32)ag_sam_atoi(s + fo[0], fl[0]);
    r.fr = (ag_sam_atoi(s + fo[1], fl[1]) & 0x10) ? 1u : 0u;
    bool star = false; u32 dot = AG_NONE;
    for (u32 i = 0; i < fl[2]; i++) { const char c = s[fo[2] + i]; if (c == '*') star = true; if (c == '.' && dot == AG_NONE) dot = i; }
    if (star) { r.tid = r.tstart = r.tend = r.tgap = r.sstart = r.send = r.sgap = r.ssize = AG_NONE; return; }   // unaligned (AG:203-209)
    const int pos = ag_sam_atoi(s + fo[3], fl[3]);
    int ins = 0, del = 0, total = 0, start = 0, end_clip = 0, lead = 1, num = 0;
    bool have_num = false;
    for (u32 i = 0; i < fl[5]; i++) {   // CIGAR state machine (AG:211-270)
        const char c = s[fo[5] + i];
        if (c >= '0' && c <= '9') { num = have_num ? num * 10 + (c - '0') : (c - '0'); have_num = true; continue; }
        const int v = have_num ? num : 0;
        if (c == 'I') { ins += v; total += v; }
        else if (c == 'D') { del += v; }
        else if (c == 'M') {
            if (r.nseg < AG_SAM_MAXSEG) { ag_seg& sg = r.seg[r.nseg]; sg.src = (u32)total; sg.dst = (u32)(pos + total + del - start - ins - 1); sg.len = (u32)v; }
            else r.err = AG_SAM_ERR_SEGS;
            r.nseg++; total += v; lead = 0;
        }
        else if (c == 'S' && lead) { start = v; total += v; lead = 0; }
        else if (c == 'S') { end_clip = v; total += v; }
        else if (c != '*') { r.err = AG_SAM_ERR_CHAR; r.bad_char = c; return; }
        else continue;  // '*' leaves the digit buffer alone (AG:263-270)
        have_num = false; num = 0;
    }
    r.sstart = (u32)start; r.send = (u32)(total - end_clip); r.sgap = (u32)ins; r.ssize = (u32)total;
    r.tid = dot != AG_NONE ? (u32)ag_sam_atoi(s + fo[2], dot) : 0u;
    r.tstart = (u32)(pos - 1);
    r.tend = r.tstart + (u32)total + (u32)del - (u32)ins;
    r.tgap = (u32)del;
}

// pair filter of loadReadAli (AG:1261): both mates aligned, >= 60 % of the read in M columns, >= 60 % of the target span not deleted;
// unsigned arithmetic cast to double, exactly as the reference writes it
AG_HD bool ag_sam_mate_pass(const ag_samline& a, double threshold) {
    return a.tid != AG_NONE && (double)(a.send - a.sstart - a.sgap) / a.ssize >= threshold && (double)(a.tend - a.tstart - a.tgap) / (a.tend - a.tstart) >= threshold;
}

// Monotone, disjoint, merged segment list equivalent to the position set updateContig (AG:763-815) builds from one record's segments:
// zero-length segments dropped, adjacent ones merged.  Returns 0 = ok, 1 = a segment runs past the read (the reference writes out of bounds:
// reported as BOWTIE ALIGNMENT ERROR), 2 = segments overlap or go backwards (cannot come from one CIGAR; the general host path decides).
AG_HD int ag_sam_normalize(const ag_seg* in, u32 n_in, u32 rlen, ag_seg* out, u32& n_out) {
    n_out = 0;
    for (u32 i = 0; i < n_in; i++) {
        const ag_seg s = in[i];
        if (!s.len) continue;
        if ((u64)s.src + s.len > rlen) return 1;
        if (n_out) {
            ag_seg& b = out[n_out - 1];
            if (s.src < b.src + b.len || s.dst < b.dst + b.len) return 2;
            if (s.src == b.src + b.len && s.dst == b.dst + b.len) { b.len += s.len; continue; }
        }
        out[n_out++] = s;
    }
    return 0;
}

// unit position of read offset 0 in the position set of one record (later segments overwrite earlier ones, AG:809-814): the key of the
// duplicate rule AG:1650-1655
AG_HD u32 ag_sam_pos_at0(const ag_seg* segs, u32 n) {
    u32 r = AG_NONE;
    for (u32 i = 0; i < n; i++) if (segs[i].len && segs[i].src == 0) r = segs[i].dst;
    return r;
}

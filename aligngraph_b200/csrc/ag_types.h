// Shared plain-data layouts of the B200 AlignGraph hot path (host <-> device).  All integer, all little-endian u32.
//
// Vocabulary follows the reference (AlignGraph.cpp, "AG:line"): a *unit* is one chromosome or one --part slice
// (AG:3395-3409); a *contiMer* is one base of a contig threaded through the unit (AG:51-62); a *node* is one KMer entry of
// the positional de Bruijn graph (AG:78-98); an *alignment* is one read-pair alignment that survived the load-time filters
// (AG:1261, AG:1650-1655).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define AG_HD __host__ __device__ __forceinline__
#define AG_HD_COLD __host__ __device__ __noinline__  /* rare paths: kept out of line so the hot loops stay small */
#else
#define AG_HD inline
#define AG_HD_COLD inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

#define AG_NONE 0xFFFFFFFFu
#define AG_EP 5  // EP, AG:39

// One M-segment of a CIGAR / one PSL block: read offsets [src, src+len) sit on unit positions [dst, dst+len)  (AG:44-49)
struct ag_seg { u32 src, dst, len; };

// One surviving read-pair alignment, 32 bytes (2 x 16 B mate tuples).  The first M segment of either mate is inline; when a
// mate has more than one, ALL its segments are in the ext array: mate 1's nseg1 entries at ext[ext_idx], then mate 2's.
struct ag_aln {
    u32 pair;       // pair index; reads are 2*pair (mate 1) and 2*pair+1 (mate 2)
    u32 flags;      // bit0: mate 1 on reverse strand (FLAG&0x10), bit1: mate 2 reverse; bits 8-15 nseg1; bits 16-23 nseg2
    u32 dst1, sl1;  // mate 1 first segment: dst ; src | len<<16
    u32 dst2, sl2;  // mate 2 first segment
    u32 ext_idx;    // first ext segment (valid when nseg1 > 1 or nseg2 > 1)
    u32 pad;
};

// Prepared alignment (device side, written by k_prep): left/right mate resolved (AG:1657-1679), touch range computed.
struct ag_alnp {
    u32 left_read;  // (read index << 1) | rc      read index = 2*pair + mate
    u32 len_nseg;   // read length | nsegL<<16 | nsegR<<24
    u32 l_dst, l_sl;
    u32 r_dst, r_sl;
    u32 ext_l;      // ext index of the LEFT mate's segments (valid when nsegL > 1)
    u32 ext_r;      // ext index of the RIGHT mate's segments (valid when nsegR > 1)
};

// contiMer as seen from a unit position (CSR over positions, push order preserved — AG:1369-1477 iterate it in order)
struct ag_cm {
    u32 cid;    // contigID  = chunk index (AG:959)
    u32 coff;   // contigOffset
    u32 chain;  // index of this contiMer in the chain-major arrays (next contiMer of the thread = chain + 1)
    u32 term;   // chain index of the terminal contiMer (nextID == -1) of this thread
};

// One contig thread: a chunk's position set threaded through the unit (AG:884-1177).  Its contiMers are the chain-major range
// [first, term]; contigOffset of chain index k is coff_first + (k - first) for k < term and coff_term for the terminal (AG:1121-1148).
// The position-ordered contiMer table (CSR + ag_cm records) is a pure function of these descriptors and chain_pos: the device builds it.
struct ag_cthread { u32 first, term, cid, coff_first, coff_term; };

// node while it is being built (k_nodes) — the founder fields never change after creation (AG:1381)
struct ag_nodeb {
    u32 cid, coff, cid0, coff0, moff;  // match fields; mate chromosome id is 0 whenever moff != NONE
    u32 cov;                           // coverage
    u32 cnt[5];                        // A C G T N   (AG:1340-1351)
    u32 sread;                         // founder k-mer string: (read index << 1) | rc
    u32 soff_len;                      // offset | len<<16  (len 0 = empty string)
    u32 succ;                          // bit j: an edge to item j of the next position was seen during the sweep
};

// final node record used by the edge sweep and the walk (position-ordered)
struct alignas(16) ag_nodew {
    u32 succ0, succ1;  // first two successors (global node index) or NONE
    u32 moff;          // chromosomeOffset0
    u32 misc;          // bits 0-7 consensus base char; bit 8 filtered (AG:1912-1915); bit 9 contigOffset != -1; bit 10 has overflow edges; bits 11-12 walk marks
};
#define AG_NW_FILTERED 0x100u
#define AG_NW_HASCONTIG 0x200u
#define AG_NW_OVF 0x400u
#define AG_NW_TRAV 0x800u    /* traversed (AG:2013); set from the start on coverage-filtered nodes */
#define AG_NW_DETOUR 0x1000u /* the walk left this node through a contiMer thread */
#define AG_NW_STOP 0x4000u    /* sequential replay: the walk that marked this node left the chain here (follow walk_next, not fnext) */
#define AG_NW_INTERIOR 0x2000u /* entered only through its unique live predecessor (forced link): never starts a walk */

struct ag_nodem { u32 cid, coff, cid0, coff0, moff; };  // match fields of a candidate / node (AG:1293-1312)
// contig-side match fields of a final node, one 16-byte record (the fifth match field, chromosomeOffset0, is ag_nodew::moff)
struct alignas(16) ag_nodec { u32 cid, coff, cid0, coff0; };

// per-walk record (one per live node that starts a walk in AG:1976-1990), compacted in scan order
struct ag_walk {
    u32 start_node;  // global node index of the start
    u32 soff;        // startOffset
    u32 eoff;        // endOffset BEFORE the tail adjustment (AG:2142-2151)
    u32 soff0;       // startOffset0 (NONE => startID0 = -1)
    u32 eoff0;       // endOffset0 BEFORE the tail adjustment (NONE => endID0 = -1)
    u32 len;         // number of bases emitted by the loop (without the tail)
    u32 last_node;   // last node marked traversed (owner of sBak, AG:2017)
    u32 flags;       // bit0 extended; bits 1-2 end mode: 0 => kMerTag -1, 1 => kMerTag -2, 2 => kMerTag 1
    u32 tail_sread;  // founder string of last_node: (read index << 1) | rc
    u32 tail_soff_len;  // offset | len<<16
};

// forced-link chain record after list ranking: from this node to the tail of its chain
struct alignas(16) ag_chain { u32 jump, tail, len, flg; };  // flg = NUMBER of nodes with contigOffset != -1 from here to the tail

// Everything a walk needs when it arrives at a chain head, in one 32-byte record: the chain's tail, length and contig flag, and the
// tail's successors / static misc bits / mate offset / detour contiMer.  One load per chain hop instead of three dependent ones (chain -> tail -> successors).
struct alignas(32) ag_hrec { u32 tail, len, flg, ts0, ts1, tmisc, tmoff, tcm; };   // tcm: index of the ONLY contiMer at the tail's position when it is not a terminal (the detour of AG:2047-2057 is possible), else NONE
struct alignas(16) ag_hdet { u32 z, first, n, len; };   // per chain head whose tail can take the contiMer detour (hrec.tcm != NONE): terminal position z of the thread, its nodes [first, first + n), contiMers stepped over

// bases a walk contributes to its contig: the loop's bases plus s[1..] of the last node when it ended in the k-mer graph (AG:2164-2168)
AG_HD u32 ag_walk_tail_len(const ag_walk& r) { u32 slen = r.tail_soff_len >> 16; return (((r.flags >> 1) & 3) != 1 && slen > 1) ? slen - 1 : 0; }


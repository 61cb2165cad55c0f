/* aligngraph_b200 — C ABI of the B200-native AlignGraph hot path.
 *
 * The reference (baoe/AlignGraph, AlignGraph/AlignGraph.cpp, cited "AG:line") has no plugin or FFI interface; its boundary for
 * this path is the five calls of the per-chromosome loop body (AG:4768-4776) and the tmp/ files they read and write:
 *
 *     loadGenome(genome, N)                       AG:287      tmp/_genome.N.fa
 *     loadContigAlignment(genome, N)              AG:1219     tmp/_contigs.fa, tmp/_contigs_genome.N.psl  -> tmp/_initial_contigs.N.fa
 *     loadReadAlignment(genome, k, iv, N, mrl)    AG:1872     tmp/_reads.fa, tmp/_reads_genome.N.bowtie
 *     extendContigs(genome, coverage, k, N)       AG:2382     -> tmp/_pre_extended_contigs.N.fa
 *     scaffoldContigs(genome, N)                  AG:2396     -> tmp/_extended_contigs.N.fa
 *
 * This library replaces them at two levels:
 *   (1) file level  — ag_run_unit_files() is a drop-in for the five calls together (same files in, same files out);
 *   (2) array level — the same work on packed host buffers (what a host that has already parsed its text would hand over):
 *       ag_set_reads / ag_begin_unit / ag_set_contimers / ag_add_alignments / ag_build / ag_extend / ag_get_text.
 * All graph construction, pruning and walking runs in CUDA kernels on one B200 per context; there is NO CPU fallback —
 * ag_create() fails when no CUDA device is present.
 *
 * Conventions: plain pointers and sizes, no C++ types; the caller owns every input buffer (they may be freed after the call
 * returns); the context owns device memory and the returned text buffers (valid until the next call on the context).  Every
 * function returns 0 on success; otherwise ag_last_error() holds the message the reference would print before exit(-1)
 * (e.g. "CANNOT OPEN FILE!", "BOWTIE ALIGNMENT ERROR", "BROKEN BOWTIE FILE").  One context per GPU; calls on different
 * contexts are thread-safe, calls on one context are not (the reference's loop is single-threaded and not re-entrant either).
 */
#ifndef ALIGNGRAPH_B200_H
#define ALIGNGRAPH_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ag_ctx ag_ctx;

typedef struct ag_params {
    int k;                /* --kMer            (AG:4701 default 5)  */
    int insert_variation; /* --insertVariation (AG:4701 default 50) */
    int coverage;         /* --coverage        (AG:4701 default 20) */
    int device;           /* CUDA device ordinal */
} ag_params;

/* One M segment of a CIGAR: read offsets [src, src+len) lie on unit positions [dst, dst+len)   (Segment, AG:44-49) */
typedef struct ag_seg_c { uint32_t src, dst, len; } ag_seg_c;

/* One read-pair alignment that passed the load-time filters (AG:1261, AG:1650-1655), 32 bytes.
 * flags: bit0 mate 1 reverse (FLAG&0x10), bit1 mate 2 reverse, bits 8-15 number of M segments of mate 1, bits 16-23 of mate 2.
 * dstX / slX: first segment of mate X (dst ; src | len<<16).  A mate with more than one segment has ALL its segments in the ext
 * array: mate 1's at ext[ext_idx], mate 2's right after (or at ext_idx when mate 1 has a single segment). */
typedef struct ag_aln_c { uint32_t pair, flags, dst1, sl1, dst2, sl2, ext_idx, pad; } ag_aln_c;

/* contiMer (ContiMer, AG:51-62) in position-CSR order; `chain` indexes the chain-major arrays (the next contiMer of the same
 * contig thread is chain+1), `term` is the chain index of the thread's terminal contiMer (nextID == -1). */
typedef struct ag_cm_c { uint32_t cid, coff, chain, term; } ag_cm_c;
/* one contig thread (a chunk's position set threaded through the unit, updateGenomeWithContig AG:884-1177): its contiMers are the
 * chain-major range [first, term]; contigOffset of chain index k = coff_first + (k - first), the terminal (AG:1121-1148) has coff_term */
typedef struct ag_cthread_c { uint32_t first, term, cid, coff_first, coff_term; } ag_cthread_c;

typedef struct ag_stats {
    /* device milliseconds (CUDA events on the context's stream), accumulated since ag_reset_stats */
    float ms_h2d, ms_prep, ms_sort, ms_nodes, ms_finalize, ms_edges, ms_components, ms_chains, ms_walk, ms_materialize, ms_d2h;
    /* host wall seconds */
    double s_parse, s_device_section, s_post;
    uint64_t n_aln, n_nodes, n_walks, n_emitted, n_keys, n_tiles, kernel_launches, h2d_bytes, d2h_bytes;
    int walk_fallback; /* 1 when the exact sequential replay (AG:2194-2202 skip rule) had to be used */
    /* GPU-side text ingestion: device ms (staging copies + kernels) and how many files took the device / the host parser */
    float ms_ingest_reads, ms_ingest_sam;
    uint64_t sam_device, sam_host, reads_device, reads_host;
    uint64_t regrows; /* sweeps repeated with a larger node table / node overflow pool / edge overflow pool */
    uint64_t reads_windowed; /* read sets ingested for the id window of the job's SAM files only */
    float ms_stage, ms_build_kernel; /* inside ms_nodes: staging gather (k_stage) and the node sweep kernel (k_build_tma / k_build) */
    float ms_select; /* emission filter + materialisation inputs on the device (fused extension path) */
} ag_stats;

int ag_create(const ag_params* params, ag_ctx** out);
void ag_destroy(ag_ctx* ctx);
const char* ag_last_error(const ag_ctx* ctx);
/* message of the last failed ag_create (no context exists then) */
const char* ag_create_error(void);

/* ---- reads: once per run, shared by all units (the reference re-reads tmp/_reads.fa for every unit, AG:1880) ------------
 * bases2: 2 bits/base (A0 C1 G2 T3), 16 bases per word, `stride2` words per read; nmask: 1 bit/base marking characters other
 * than upper-case ACGT (counted as N, AG:1349), `stridem` words per read; reads 2p and 2p+1 are the mates of pair p and have
 * pair_len[p] bases each (AG:3454). */
int ag_set_reads(ag_ctx* ctx, const uint32_t* bases2, const uint32_t* nmask, const uint16_t* pair_len, uint64_t n_pairs, uint32_t stride2, uint32_t stridem);
/* same layout, buffers already in device memory of this context's GPU (target of the NCCL broadcast) */
int ag_set_reads_device(ag_ctx* ctx, const uint32_t* d_bases2, const uint32_t* d_nmask, const uint16_t* d_pair_len, uint64_t n_pairs, uint32_t stride2, uint32_t stridem);
/* original characters of the masked bases (only needed to print them in contig tails, AG:2167): keys = read*65536 + offset, sorted */
int ag_set_read_exceptions(ag_ctx* ctx, const uint64_t* keys, const char* chars, uint64_t n);
/* parse tmp/_reads.fa (AG:361-404) and upload */
int ag_load_reads_fasta(ag_ctx* ctx, const char* path);
/* the same, but only the reads that the given units can reference are made resident: the pair-id window spanned by their SAM files
 * (tmp/_reads_genome.N.bowtie), found in the read file by bisection (ids are the pair indices, AG:3455-3471); what ag_run_job_files does */
int ag_load_reads_for_units(ag_ctx* ctx, const char* reads_fa, const char* tmp_dir, const int* units, int n_units);
/* packed host copy held by the context (after ag_load_reads_fasta) — lets a launcher broadcast it */
int ag_get_reads(ag_ctx* ctx, const uint32_t** bases2, const uint32_t** nmask, const uint16_t** pair_len, uint64_t* n_pairs, uint32_t* stride2, uint32_t* stridem);

/* one read set for every context of a run (one context per GPU): ctxs[0] holds it (ag_load_reads_fasta / ag_set_reads*); the others share
 * its host copy and receive the packed device buffers through ONE broadcast — ncclBroadcast over NVLink when libnccl is loadable and the
 * devices are distinct, cudaMemcpyPeerAsync otherwise (replaces the per-chromosome re-read of tmp/_reads.fa, AG:1880) */
typedef struct ag_bcast_info { double seconds; uint64_t bytes; int nccl; } ag_bcast_info;
int ag_broadcast_reads(ag_ctx** ctxs, int n_ctx, ag_bcast_info* info /* may be NULL */);

/* ---- array level -------------------------------------------------------------------------------------------------------------- */
int ag_begin_unit(ag_ctx* ctx, int unit_id, const char* ref_bases, uint32_t n_ref);                       /* loadGenome, AG:287 */
int ag_set_contimers(ag_ctx* ctx, const uint32_t* cm_start /* n_ref+n_tail+1 */, const ag_cm_c* cm, uint32_t n_cm,
                     const uint32_t* chain_pos, const char* chain_base, const char* tail_bases, uint32_t n_tail); /* result of loadContigAlignment, AG:1219 */
/* the same result of loadContigAlignment in compact form: contig threads in push order + the chain-major position / base arrays; the
 * position-ordered table (cm_start, cm) is then derived on the device (20 bytes per contiMer less to upload) */
int ag_set_contig_threads(ag_ctx* ctx, const ag_cthread_c* threads, uint32_t n_threads, const uint32_t* chain_pos, const char* chain_base, uint32_t n_cm,
                          const char* tail_bases, uint32_t n_tail);
int ag_add_alignments(ag_ctx* ctx, const ag_aln_c* aln, uint64_t n, const ag_seg_c* ext, uint64_t n_ext); /* parsing half of loadReadAlignment, AG:1872 */
int ag_build(ag_ctx* ctx);   /* graph half of loadReadAlignment: updateGenomeWithRead / updateKMer, AG:1635-1870, AG:1353-1624 */
int ag_extend(ag_ctx* ctx);  /* extendContigs + scaffoldContigs, AG:2382, AG:2396 */
/* ag_build + ag_extend as one step: every kernel from the uploaded unit to the walk records is queued without a host round trip and the
 * host synchronises once (ag_build alone synchronises at its end so that it can report its own errors) */
int ag_process(ag_ctx* ctx);
/* which: 0 = tmp/_initial_contigs.N.fa (file level only), 1 = tmp/_pre_extended_contigs.N.fa, 2 = tmp/_extended_contigs.N.fa */
int ag_get_text(ag_ctx* ctx, int which, const char** text, uint64_t* len);

/* ---- file level ------------------------------------------------------------------------------------------------------------------- */
/* host half only: parse the unit's files into the staged arrays (then ag_build / ag_extend / ag_write_unit_files) */
int ag_prepare_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id);
int ag_write_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id);
/* the five calls of AG:4768-4776 for unit N */
int ag_run_unit_files(ag_ctx* ctx, const char* tmp_dir, int unit_id);
/* units [first_unit, first_unit + n_units) through ag_run_unit_files, one worker per context (one context per GPU), with `prefetch` host
 * threads parsing units ahead of the GPUs.  `done` (may be NULL) is called after every unit, serialised, possibly out of unit order. */
int ag_run_units_files(ag_ctx** ctxs, int n_ctx, const char* tmp_dir, int first_unit, int n_units, int prefetch,
                       void (*done)(int unit, int rc, const char* error, void* user), void* user);
/* the whole hot loop (AG:4765-4783) for a list of units, read set included: tmp/_reads.fa (reads_fa; NULL = the contexts already hold the
 * reads) is ingested on ctxs[0]'s GPU and broadcast to the other contexts while the preparer threads already parse the first units */
int ag_run_job_files(ag_ctx** ctxs, int n_ctx, const char* tmp_dir, const char* reads_fa, const int* units, int n_units, int prefetch,
                     void (*done)(int unit, int rc, const char* error, void* user), void* user);
/* staged arrays of the current unit (valid until the next ag_begin_unit / ag_prepare_unit_files) */
typedef struct ag_unit_view {
    const char* ref; uint32_t n_ref, n_tail;
    const uint32_t* cm_start; const ag_cm_c* cm; uint32_t n_cm; const uint32_t* chain_pos; const char* chain_base;
    const ag_aln_c* aln; uint64_t n_aln; const ag_seg_c* ext; uint64_t n_ext;
    const ag_cthread_c* threads; uint32_t n_threads;   /* contig threads (empty when the unit was staged through ag_set_contimers) */
} ag_unit_view;
int ag_get_unit(ag_ctx* ctx, ag_unit_view* out);

/* ---- introspection ------------------------------------------------------------------------------------------------------------------- */
int ag_get_stats(ag_ctx* ctx, ag_stats* out);
int ag_reset_stats(ag_ctx* ctx);
/* node table after ag_build as text, one node per line (tests): pos item cov A C G T N cid coff cid0 coff0 mid moff [s] pos:item...
 * The per-node coverage / base counters are only kept when ag_keep_node_counts(ctx, 1) was called before ag_build (24 B per node). */
int ag_keep_node_counts(ag_ctx* ctx, int on);
int ag_dump_nodes_text(ag_ctx* ctx, const char** text, uint64_t* len);
void* ag_cuda_stream(ag_ctx* ctx);
/* ---- benchmark support ---------------------------------------------------------------------------------------------------------
 * ag_build() uploads the staged unit arrays only when they changed since the last upload; ag_invalidate_device_inputs() forces the
 * next ag_build() to copy them again (end-to-end timing), ag_reupload_reads() repeats the reads' host->device copy.
 * ag_timer_start/stop bracket a region with CUDA events on the context's stream. */
int ag_invalidate_device_inputs(ag_ctx* ctx);
int ag_reupload_reads(ag_ctx* ctx);
/* page-lock the staged host arrays (reads + current unit) so that the copies above run from pinned memory */
int ag_pin_staged(ag_ctx* ctx);
/* input normalisation that --resume re-runs (formalizeInput(contigs) + formalizeGenome, AG:4757-4758): writes tmp/_contigs.fa,
 * tmp/_chaff.fa, tmp/_genome.fa and tmp/_genome.N.fa; returns the number of units */
int ag_formalize_inputs(ag_ctx* ctx, const char* contig_fa, const char* genome_fa, const char* tmp_dir, int part, int* n_units);
/* the aligner call inside refinement() (AG:2957-2983: pblat <db> <query> -noHead <out> -fastMap) for hosts without BLAT: ungapped seed-and-verify
 * containment search of every query (the truncated initial contigs, <= 20 kbp) in every database sequence (the extended contigs), either
 * strand — exact 24-mer seeds, full-length comparison on the GPU (>= 90 % identity), local X-drop alignments for queries that do not fit
 * anywhere — written as single-block PSL lines */
int ag_containment_search_files(ag_ctx* ctx, const char* db_fa, const char* query_fa, const char* out_psl);
/* removeMisassembly(file, distanceLow, distanceHigh, id, coverage, fastMap), AG:4281-4297, for one output file (`id` = "extended" or
 * "remaining"): formalizes `file` into tmp/_<id>_contigs.fa, calls `align(id, user)` (non-zero = ok) which must run the reference's aligner
 * commands (AG:3825-3849: bowtie2 reads -> contigs into tmp/_reads_<id>_contigs.bowtie, BLAT contigs -> genome into tmp/_<id>_contigs_genome.psl),
 * piles the read alignments up into a per-base coverage on the GPU, and writes corrected_<file> in the current directory */
int ag_remove_misassembly_file(ag_ctx* ctx, const char* file, const char* id, int coverage, const char* tmp_dir, int (*align)(const char* id, void* user), void* user);
/* tuning / test hooks: "reads_window" (0 = ag_run_job_files always makes the whole read set resident), "host_parse" (1 = SAM and reads text parsed by the host parsers instead of the device kernels), "node_cap" / "ovf_cap" /
 * "eovf_cap" / "key_cap" / "cand_cap" / "hwalk_cap" (initial capacities of the node table, the node overflow pool, the edge overflow pool, the
 * tile-key buffers, the start-candidate arrays and the host walk-record buffer; small values force the grow-and-redo path), "rank_rounds"
 * (global list-ranking rounds queued per step), "bases_cap" (initial capacity of the materialised-bases buffer), "fused_extend" (0 = emission filter on the host between walk and
 * materialisation: two synchronisations per step), "tma" (0 = node sweep with per-thread staging, k_build, instead of the bulk-async staged k_build_tma) */
int ag_set_option(ag_ctx* ctx, const char* name, long value);
int ag_timer_start(ag_ctx* ctx);
int ag_timer_stop(ag_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif

// B200 (sm_100a) device pipeline of the AlignGraph hot path: positional de Bruijn graph build (AG:1635-1870, AG:1353-1624),
// coverage filter (AG:1904-1918) and extension walk (AG:1954-2204).  Integer / indexing work, HBM- and latency-bound: no tensor
// cores.  See DESIGN.md §3-§5 for the formulation, the HBM layout and the per-kernel algorithmic bytes.
//
//   k_scan_onepass   every prefix sum of the step: chained scan with decoupled look-back, one launch
//   k_chain_expand / k_cm_count / k_cm_fill / k_cm_sort / k_cm1   contig-thread descriptors -> chain arrays -> contiMer table (CSR + summary)  (AG:884-1177)
//   k_prep        1 thread / alignment   resolve left mate, touch range, tile count, clean / linear bits                (AG:1657-1679)
//   k_keys        1 thread / alignment   emit (tile, alignment) keys in alignment order
//   k_rs_hist / k_rs_scatter   stable LSD radix sort, 5-bit digits: alignments bucketed by 248-position tile, order preserved
//   k_stage       8 lanes / tile key     prepared header + oriented 4-bit codes of the left mate -> one contiguous record per key
//   k_build_tma   1 CTA / tile, 1 thread / position (+ 1 halo lane per warp)   records staged by cp.async.bulk + mbarrier; ordered
//                 first-compatible clustering (AG:1353-1587), coverage filter + consensus base (AG:1904-1918, 1944-1952) and the
//                 common-case edges (AG:1590-1623) in one sweep   (k_build: the same sweep with per-thread staging, reads > ~500 bp)
//   k_posfix / k_succ   tile blocks -> position order, successor-item bits -> node indices
//   k_edges       generic edge sweep, flagged tiles only                                                          (AG:1590-1623)
//   k_indeg / k_links_rank_local / k_rank / k_cand_scatter / k_hrec   forced-link chains, start candidates, hop + detour records
//   k_uf_*        union-find over chain tails: independent walk components
//   k_walk_components (1 warp / component) | k_walk_sequential (skip rule)   exact replay of the greedy walk    (AG:1972-2204)
//   k_walk_compact, k_sel_*, k_excl_max_scan   walk records in scan order, emission filter (AG:2176-2189) as a prefix maximum
//   k_mat_*       base strings of the emitted walks; k_to_host16 copies to page-locked memory; k_occupancy  bitmap for the scaffold gap test (AG:2428)
//   ag_ingest.cuh: k_nl_*, k_rd_*, k_sam_*   text ingestion (reads FASTA, SAM), k_cov_marks, k_verify_placements
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound with dlopen (ag_device_broadcast_reads)
#include <atomic>
#include <memory>
#include <thread>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include "ag_core.h"
#include "ag_device.cuh"
#include "ag_host.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw AgError{std::string(#x) + ": " + cudaGetErrorString(e_)}; } while (0)

namespace {

struct PinnedBuf {  // page-locked host staging buffer for device->host results
    char* p = nullptr; size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t nc = n + n / 8 + 4096;
        CK(cudaHostAlloc((void**)&p, nc, cudaHostAllocDefault));
        cap = nc;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

template <class T> struct DBuf {
    T* p = nullptr; size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t nc = n + n / 16 + 256;
        CK(cudaMalloc((void**)&p, nc * sizeof(T)));
        cap = nc;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ---------------------------------------------------------------------------------------------------------------------------
// exclusive scan (u32), out has n + 1 entries (out[n] = total)
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* smem /* 32 */, u32& total) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u32 x = v;
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nw ? smem[lane] : 0, s = w;
        for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += y; }
        smem[lane] = s - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = s;
    }
    __syncthreads();
    u32 r = x - v + smem[warp];
    total = smem[32];
    __syncthreads();
    return r;
}
// exclusive prefix MAXIMUM over u32 (identity 0), one element per thread of a single CTA with a running carry: out[i] = max(in[0 .. i-1]).
// The emission filter of extdContigs1 (AG:2176-2189) over the walk records is exactly this (fused extension path); a few 10^5 elements at most.
__global__ void __launch_bounds__(1024) k_excl_max_scan(const u32* __restrict__ in, u32* __restrict__ out, const u32* __restrict__ n_ptr, u32 cap) {
    __shared__ u32 sw[32];
    __shared__ u32 carry;
    const u32 n = min(*n_ptr, cap), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < n ? in[i] : 0u;
        u32 x = v;   // inclusive warp max
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= (u32)o) x = max(x, y); }
        if (lane == 31) sw[warp] = x;
        __syncthreads();
        u32 pre = carry;   // maximum of everything before this warp
        for (u32 w = 0; w < warp; w++) pre = max(pre, sw[w]);
        const u32 up = __shfl_up_sync(0xFFFFFFFFu, x, 1);
        if (i < n) out[i] = lane ? max(pre, up) : pre;
        __syncthreads();
        if (threadIdx.x == 1023) carry = max(pre, x);
        __syncthreads();
    }
}

__global__ void k_scan_sums(const u32* __restrict__ in, u32* __restrict__ sums, size_t n) {
    __shared__ u32 sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_B + (size_t)threadIdx.x * SCAN_I;
    u32 s = 0;
    for (int i = 0; i < SCAN_I; i++) if (base + i < n) s += in[base + i];
    u32 total; block_excl_scan(s, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void k_scan_apply(const u32* __restrict__ in, u32* __restrict__ out, const u32* __restrict__ offs, size_t n, int write_total) {
    __shared__ u32 sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_B + (size_t)threadIdx.x * SCAN_I;
    u32 v[SCAN_I], s = 0;
    for (int i = 0; i < SCAN_I; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    u32 total; u32 ex = block_excl_scan(s, sm, total) + (offs ? offs[blockIdx.x] : 0);
    for (int i = 0; i < SCAN_I; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_T - 1) out[n] = ex;
}

// Single-pass scan: every CTA scans its 2048 elements, publishes {epoch, flag, value} as ONE 64-bit word (flag 1 = the block's own sum,
// 2 = inclusive prefix up to and including the block) and warp 0 looks back over its predecessors' words, 32 at a time, until it meets an inclusive
// prefix (chained scan with decoupled look-back).  One launch and one read + one write of the data instead of three launches and two reads.  The
// epoch (a per-call number from the host, 30 bits) makes words left by earlier calls read as "not ready", so the state array is never cleared.
// CTAs are dispatched in blockIdx order, so every predecessor a CTA waits for is resident or finished.
__device__ __forceinline__ u64 ld_state(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_state(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
// the scan's operator: + (identity 0), or max over u32 (identity 0: the emission filter's prefix maximum)
template <bool MAX> __device__ __forceinline__ u32 scan_op(u32 a, u32 b) { return MAX ? max(a, b) : a + b; }
// exclusive block scan of one value per thread (SCAN_T threads), total = the block's reduction
template <bool MAX> __device__ __forceinline__ u32 block_excl_scan_op(u32 v, u32* smem /* 33 */, u32& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u32 x = v;   // inclusive inside the warp
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x = scan_op<MAX>(x, y); }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const u32 w = lane < nw ? smem[lane] : 0u;
        u32 sx = w;
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, sx, o); if (lane >= o) sx = scan_op<MAX>(sx, y); }
        const u32 up = __shfl_up_sync(0xFFFFFFFFu, sx, 1);
        smem[lane] = lane ? up : 0u;   // exclusive over the warps
        if (lane == 31) smem[32] = sx;
    }
    __syncthreads();
    const u32 upx = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    const u32 r = scan_op<MAX>(smem[warp], lane ? upx : 0u);
    total = smem[32];
    __syncthreads();
    return r;
}
// n_ptr (optional): the number of elements is only known on the device (at most n_cap, which sizes the grid): blocks behind it leave at once, the
// total still goes to out[n_cap] and the entries between are not written.
template <bool MAX>
__global__ void __launch_bounds__(SCAN_T) k_scan_onepass(const u32* in, u32* out, u64* __restrict__ state, size_t n_cap, const u32* __restrict__ n_ptr, u32 epoch, int write_total) {
    __shared__ u32 sm[33];
    __shared__ u32 s_prefix;
    const u32 b = blockIdx.x, lane = threadIdx.x & 31;
    const size_t n = n_ptr ? min((size_t)*n_ptr, n_cap) : n_cap;
    const u32 last_b = n ? (u32)((n - 1) / SCAN_B) : 0u;
    if (b > last_b) return;
    const size_t base = (size_t)b * SCAN_B + (size_t)threadIdx.x * SCAN_I;
    u32 v[SCAN_I], s = 0;
    if (base + SCAN_I <= n && (((size_t)in) & 15) == 0) {
        const uint4 a = *(const uint4*)(in + base), c = *(const uint4*)(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
        for (int i = 0; i < SCAN_I; i++) v[i] = (base + i < n) ? in[base + i] : 0;
    }
    for (int i = 0; i < SCAN_I; i++) s = scan_op<MAX>(s, v[i]);
    u32 total; u32 ex = block_excl_scan_op<MAX>(s, sm, total);
    const u64 tag = (u64)epoch << 34;
    if (threadIdx.x == 0) { s_prefix = 0; st_state(state + b, tag | ((u64)(b == 0 ? 2 : 1) << 32) | total); }
    if (b > 0 && threadIdx.x < 32) {
        u32 run = 0;
        for (long long j = (long long)b - 1;; j -= 32) {
            const long long idx = j - lane;                      // lane 0 = the nearest predecessor
            u64 w; u32 flag;
            do {
                w = idx >= 0 ? ld_state(state + idx) : (tag | (2ull << 32));
                flag = (w >> 34) == epoch ? (u32)(w >> 32) & 3u : 0u;
            } while (__any_sync(0xFFFFFFFFu, flag == 0));
            const u32 incl = __ballot_sync(0xFFFFFFFFu, flag == 2);
            const u32 first = incl ? (u32)__ffs(incl) - 1 : 31u;
            u32 x = lane <= first ? (u32)w : 0u;
            for (int o = 16; o; o >>= 1) x = scan_op<MAX>(x, __shfl_xor_sync(0xFFFFFFFFu, x, o));
            run = scan_op<MAX>(run, x);
            if (incl) break;
        }
        if (lane == 0) { s_prefix = run; st_state(state + b, tag | (2ull << 32) | scan_op<MAX>(run, total)); }
    }
    __syncthreads();
    ex = scan_op<MAX>(ex, s_prefix);
    if (base + SCAN_I <= n && (((size_t)out) & 15) == 0) {
        uint4 a, c;
        a.x = ex; ex = scan_op<MAX>(ex, v[0]); a.y = ex; ex = scan_op<MAX>(ex, v[1]); a.z = ex; ex = scan_op<MAX>(ex, v[2]); a.w = ex; ex = scan_op<MAX>(ex, v[3]);
        c.x = ex; ex = scan_op<MAX>(ex, v[4]); c.y = ex; ex = scan_op<MAX>(ex, v[5]); c.z = ex; ex = scan_op<MAX>(ex, v[6]); c.w = ex; ex = scan_op<MAX>(ex, v[7]);
        *(uint4*)(out + base) = a; *(uint4*)(out + base + 4) = c;
    } else {
        for (int i = 0; i < SCAN_I; i++) { if (base + i < n) out[base + i] = ex; ex = scan_op<MAX>(ex, v[i]); }
    }
    if (write_total && b == last_b && threadIdx.x == SCAN_T - 1) out[n_cap] = ex;
}

struct Scanner {
    DBuf<u32> lvl[4];
    DBuf<u64> state;
    u32 epoch = 0;
    bool one_pass = getenv("AG_SCAN_TWOPASS") == nullptr;   // (A/B switch for measurements)
    u64* launches = nullptr;
    // n_ptr (one-pass kernel only; callers check one_pass): device-side element count <= n
    // exclusive prefix MAXIMUM (identity 0) of in[0, min(*n_ptr, n)): out[i] = max(in[0 .. i-1]); one launch, any size
    void run_max(const u32* in, u32* out, size_t n, const u32* n_ptr, cudaStream_t st) {
        if (n == 0) return;
        const size_t nb = (n + SCAN_B - 1) / SCAN_B;
        if (nb > state.cap) { state.ensure(nb); CK(cudaMemsetAsync(state.p, 0, state.cap * sizeof(u64), st)); }
        epoch = (epoch + 1) & 0x3FFFFFFFu; if (!epoch) epoch = 1;
        k_scan_onepass<true><<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, state.p, n, n_ptr, epoch, 0);
        if (launches) ++*launches;
    }
    void run(const u32* in, u32* out, size_t n, cudaStream_t st, int depth = 0, int write_total = 1, const u32* n_ptr = nullptr) {
        if (n == 0) { if (write_total) CK(cudaMemsetAsync(out, 0, sizeof(u32), st)); return; }
        size_t nb = (n + SCAN_B - 1) / SCAN_B;
        if (one_pass) {
            if (nb > state.cap) { state.ensure(nb); CK(cudaMemsetAsync(state.p, 0, state.cap * sizeof(u64), st)); }
            epoch = (epoch + 1) & 0x3FFFFFFFu; if (!epoch) epoch = 1;
            k_scan_onepass<false><<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, state.p, n, n_ptr, epoch, write_total);
            if (launches) ++*launches;
            return;
        }
        if (nb == 1) { k_scan_apply<<<1, SCAN_T, 0, st>>>(in, out, nullptr, n, write_total); if (launches) ++*launches; return; }
        if (depth >= 4) throw AgError{"scan depth"};
        lvl[depth].ensure(nb + 1);
        k_scan_sums<<<(unsigned)nb, SCAN_T, 0, st>>>(in, lvl[depth].p, n);
        if (launches) ++*launches;
        run(lvl[depth].p, lvl[depth].p, nb, st, depth + 1, 0);
        k_scan_apply<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, lvl[depth].p, n, write_total);
        if (launches) ++*launches;
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key = tile, val = alignment index), RS_BITS bits per pass
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int RS_T = 256, RS_I = 8, RS_B = RS_T * RS_I, RS_BITS = 5, RS_BINS = 1 << RS_BITS;

__global__ void k_rs_hist(const u32* __restrict__ keys, u32* __restrict__ hist, const u32* __restrict__ n_ptr, u32 cap, int shift, unsigned nb) {
    __shared__ u32 h[RS_BINS];
    const size_t n = min(*n_ptr, cap);   // (more keys than the buffers hold: E_KEY_CAP is set, the step is repeated)
    if (threadIdx.x < RS_BINS) h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_B;
    for (int i = threadIdx.x; i < RS_B; i += RS_T) if (base + i < n) atomicAdd(&h[(keys[base + i] >> shift) & (RS_BINS - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < RS_BINS) hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void k_rs_scatter(const u32* __restrict__ keys, const u32* __restrict__ vals, u32* __restrict__ okeys, u32* __restrict__ ovals,
                             const u32* __restrict__ hist_scanned, const u32* __restrict__ n_ptr, u32 cap, int shift, unsigned nb) {
    const size_t n = min(*n_ptr, cap);
    // counters in flat (digit, thread) order, f = digit * RS_T + thread, stored skewed by one word per 32 so that both access patterns —
    // [digit][thread] while counting / ranking and "thread t owns flat entries [RS_BINS t, RS_BINS t + RS_BINS)" in the scan — are
    // free of bank conflicts
    constexpr int NF = RS_BINS * RS_T;
    __shared__ u32 cnt[NF + NF / 32];
    __shared__ u32 sm[33];
    __shared__ u32 bin_start[RS_BINS + 1];
    auto at = [&](int f) -> u32& { return cnt[f + (f >> 5)]; };
    const int t = threadIdx.x;
    size_t base = (size_t)blockIdx.x * RS_B + (size_t)t * RS_I;
    u32 k[RS_I], v[RS_I];
    for (int d = 0; d < RS_BINS; d++) at(d * RS_T + t) = 0;
    for (int i = 0; i < RS_I; i++)
        if (base + i < n) { k[i] = keys[base + i]; v[i] = vals[base + i]; at((int)((k[i] >> shift) & (RS_BINS - 1)) * RS_T + t)++; }
    __syncthreads();
    u32 s = 0;
    for (int j = 0; j < RS_BINS; j++) s += at(t * RS_BINS + j);
    u32 total; u32 ex = block_excl_scan(s, sm, total);
    for (int j = 0; j < RS_BINS; j++) { const u32 c = at(t * RS_BINS + j); at(t * RS_BINS + j) = ex; ex += c; }
    __syncthreads();
    if (t < RS_BINS) bin_start[t] = at(t * RS_T);
    if (t == 0) bin_start[RS_BINS] = total;
    __syncthreads();
    for (int i = 0; i < RS_I; i++)
        if (base + i < n) {
            u32 d = (k[i] >> shift) & (RS_BINS - 1);
            u32 r = at((int)d * RS_T + t)++;
            size_t dst = (size_t)hist_scanned[(size_t)d * nb + blockIdx.x] + (r - bin_start[d]);
            okeys[dst] = k[i]; ovals[dst] = v[i];
        }
}

}  // namespace
#include "ag_ingest.cuh"
namespace {

// ---------------------------------------------------------------------------------------------------------------------------
// device view of one unit
// ---------------------------------------------------------------------------------------------------------------------------
// error word of a step (bits, set with atomicOr; read once at the step's single synchronisation point)
enum : int { E_OVF = 1, E_BAD_ALN = 2, E_NODE_CAP = 4, E_EDGE_OVF = 8, E_WALK_CAP = 16, E_MAT_CAP = 32, E_CM = 64, E_KEY_CAP = 128, E_CAND_CAP = 256, E_HWALK_CAP = 512, E_RANK_MORE = 1024, E_BASES_CAP = 2048 };
// fixed launch geometry of the loops over nodes / candidates whose trip count only the device knows
constexpr unsigned GS_BLOCKS = 148 * 8, GS_T = 256;
constexpr int E_FATAL = E_OVF | E_NODE_CAP | E_EDGE_OVF | E_CM | E_KEY_CAP | E_CAND_CAP | E_RANK_MORE | E_HWALK_CAP | E_BASES_CAP;   // the step is repeated with larger capacities: later kernels skip their work
#define AG_BAIL(d) do { if (*(volatile int*)(d).err & E_FATAL) return; } while (0)
#define AG_FOR_N(v, n) for (u32 v = blockIdx.x * blockDim.x + threadIdx.x, n_ = (n), s_ = gridDim.x * blockDim.x; v < n_; v += s_)

struct DevView {
    ag_reads reads;
    const unsigned char* ref; u32 n_ref, n_pos;
    ag_cmtab cmt; const ag_cm1* cm1; const u32* many_prefix; const u32* lin_prefix; const u32* chain_pos; const unsigned char* chain_base;
    const ag_aln* aln; u32 n_aln; const ag_seg* ext;
    ag_alnp* alnp; ag_fast* fast; u32* ntiles; u32* key_off;
    u32* keys; u32* vals; u32* tile_cnt; u32* tile_start; u32 n_tiles;
    ag_ovfpool ovf;
    u32* tile_flag; u32* tile_base; u32* tile_nodes; u32* tile_prefix; u32* pool_count; u32 node_cap;
    u32* pos_pool; u32* pos_node;
    ag_nodec* pool_c; ag_nodew* pool_w; u32* pool_sref; u32* pool_pos; u32* pool_cc;      // node records in sweep order (tile blocks in arbitrary order)
    ag_nodec* node_c; ag_nodew* node_w; u32* node_sref; u32* node_pos; u32* node_cc;      // the same in position order (k_succ)
    u32* eovf_head; u32* eovf_target; u32* eovf_next; u32* eovf_count; u32 eovf_cap;
    u32* walk_next; u32* parent; u32* cmin; u32* cmax;
    u32* cand_rank; u32* cand_node; u32* cand_label;
    unsigned char* pos_term; u32* indeg; u32* fnext; u32* fprev; u32* msuf; u32* mnode; ag_chain* chain_a; ag_chain* chain_b; const ag_chain* chain; ag_hrec* hrec; ag_hdet* hdet; int* changed;
    ag_walk* walks; ag_walk* walks_sorted; u32* walk_count; u32 walk_cap; u32* walk_used;
    int* err;
    int k, iv, coverage;
    const int* err_load;                   // error word of the unit upload (contig thread kernels)
    const u32* nk_ptr; u32 key_cap;        // number of tile keys (device), capacity of the key buffers
    const u32* nn_ptr;                     // number of nodes (device) = pool_count
    const u32* ncand_ptr; u32 cand_cap;    // number of walk start candidates (device), capacity of the per-candidate arrays
    ag_walk* walks_host; u32 hwalk_cap;    // page-locked host buffer the compacted walk records are written to
    u32 bases_cap;                         // capacity of the materialised base buffer (fused extension path)
    u32 rw;  // words per staged read (stride2 + stridem) when the tile sweeps keep the chunk's reads in shared memory, else 0
};

// ---------------------------------------------------------------------------------------------------------------------------
// k_prep / k_keys
// ---------------------------------------------------------------------------------------------------------------------------
// 16-byte vector copies of whole records (sizes are multiples of 16 and the arrays come from cudaMalloc): a 48-byte record leaves as three
// 128-bit stores instead of twelve 32-bit ones
template <class T> __device__ __forceinline__ void st_rec16(T* dst, const T& v) {
    static_assert(sizeof(T) % 16 == 0, "record size");
    uint4 t[sizeof(T) / 16]; memcpy(t, &v, sizeof(T));
    for (unsigned k = 0; k < sizeof(T) / 16; k++) reinterpret_cast<uint4*>(dst)[k] = t[k];
}
template <class T> __device__ __forceinline__ T ld_rec16(const T* src) {
    static_assert(sizeof(T) % 16 == 0, "record size");
    uint4 t[sizeof(T) / 16];
    for (unsigned k = 0; k < sizeof(T) / 16; k++) t[k] = reinterpret_cast<const uint4*>(src)[k];
    T v; memcpy(&v, t, sizeof(T)); return v;
}
__global__ void k_prep(DevView d) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_aln) return;
    ag_aln a = ld_rec16(d.aln + i);
    u32 len = d.reads.len[a.pair];
    ag_prep_out o = ag_prep(a, d.ext, len, (u32)d.k);
    // every aligned position of either mate must lie inside the unit (the reference indexes genome[0] with them, AG:1369)
    ag_segv L, R; ag_alnp_segs(o.p, d.ext, L, R);
    bool bad = false;
    for (u32 j = 0; j < L.n; j++) { ag_seg s = L.get(j); if (s.dst + s.len > d.n_ref || s.src + s.len > len) bad = true; }
    for (u32 j = 0; j < R.n; j++) { ag_seg s = R.get(j); if (s.dst + s.len > d.n_ref || s.src + s.len > len) bad = true; }
    if (o.any && o.lo + o.span >= d.n_ref) bad = true;
    if (bad) { atomicOr(d.err, E_BAD_ALN); o.any = 0; }
    ag_fast f = ag_fast_prep(o.p, o.lo, o.span, i);
    if (o.any) ag_fast_classify(f, o.p, d.many_prefix, d.lin_prefix, d.cm1);
    st_rec16(d.alnp + i, o.p); st_rec16(d.fast + i, f);
    u32 nt = 0;
    if (o.any) { u32 t0, t1; ag_tile_range(o.lo, o.lo + o.span, d.n_tiles, t0, t1); nt = t1 - t0 + 1; }
    d.ntiles[i] = nt;
}

__global__ void k_keys(DevView d) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_aln) return;
    u32 n = d.ntiles[i];
    if (!n) return;
    const ag_fast f = d.fast[i];
    u32 t0, t1; ag_tile_range(f.lo, f.lo + f.span, d.n_tiles, t0, t1);
    u32 off = d.key_off[i];
    if ((u64)off + n > d.key_cap) { atomicOr(d.err, E_KEY_CAP); return; }
    for (u32 j = 0; j < n; j++) { d.keys[off + j] = t0 + j; d.vals[off + j] = i; atomicAdd(&d.tile_cnt[t0 + j], 1u); }
}

// non-ACGT bit plane of the reads from its list of set bits (key = read * 65536 + offset)
__global__ void k_nmask_scatter(const u64* __restrict__ keys, u64 n, u32* nmask, u32 stridem, u64 n_reads) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 read = keys[i] >> 16; const u32 off = (u32)(keys[i] & 0xFFFFu);
    if (read < n_reads && (off >> 5) < stridem) atomicOr(&nmask[read * stridem + (off >> 5)], 1u << (off & 31));
}

// ---------------------------------------------------------------------------------------------------------------------------
// contiMer table from the contig threads (loadContigAlignment's product, AG:884-1177, re-derived on the device instead of being
// uploaded: 20 bytes per contiMer stay off the PCIe link).  Position p's list holds the contiMers with chain_pos == p in push order,
// i.e. by increasing chain index: count -> scan -> fill in arrival order -> per-position insertion sort by chain index.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void k_cm_count(const u32* __restrict__ chain_pos, u32 n_cm, u32 n_pos, u32* cnt, int* err) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cm) return;
    u32 p = chain_pos[k];
    if (p >= n_pos) { atomicOr(err, E_CM); return; }
    atomicAdd(&cnt[p], 1u);
}
__global__ void k_cm_fill(const u32* __restrict__ chain_pos, u32 n_cm, const ag_cthread* __restrict__ th, u32 n_th, const u32* __restrict__ cm_start, u32* fill, ag_cm* cm, int* err) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cm) return;
    u32 lo = 0, hi = n_th;   // last thread with first <= k
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (th[mid].first <= k) lo = mid; else hi = mid; }
    const ag_cthread t = th[lo];
    if (n_th == 0 || k < t.first || k > t.term) { atomicOr(err, E_CM); return; }
    const u32 p = chain_pos[k];
    ag_cm m; m.cid = t.cid; m.coff = k < t.term ? t.coff_first + (k - t.first) : t.coff_term; m.chain = k; m.term = t.term;
    cm[cm_start[p] + atomicAdd(&fill[p], 1u)] = m;
}
__global__ void k_cm_sort(const u32* __restrict__ cm_start, u32 n_pos, ag_cm* cm) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    const u32 a = cm_start[p], n = cm_start[p + 1] - a;
    for (u32 i = 1; i < n; i++) {
        ag_cm x = cm[a + i]; u32 j = i;
        while (j > 0 && cm[a + j - 1].chain > x.chain) { cm[a + j] = cm[a + j - 1]; j--; }
        cm[a + j] = x;
    }
}

// chain-major contiMer arrays from the run-space contig threads (ag_thread_contigs_runs): chain index k of thread t is base F + (k - first)
// of the oriented chunk; it sits on the unit position its run maps it to, or — an inserted base between two runs — in the tail behind the
// unit; the thread's last contiMer is the terminal and carries the UNIT's base (AG:1121-1148).  One thread per contiMer.
__global__ void k_chain_expand(const ag_cdesc* __restrict__ desc, u32 n_desc, const ag_crun* __restrict__ runs, const char* __restrict__ blob, const unsigned char* __restrict__ ref,
                               u32 n_ref, u32 n_cm, u32* __restrict__ chain_pos, unsigned char* __restrict__ chain_base) {
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cm) return;
    u32 lo = 0, hi = n_desc;   // last thread with first <= k
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (desc[mid].first <= k) lo = mid; else hi = mid; }
    const ag_cdesc d = desc[lo];
    const ag_crun* r = runs + d.run0;
    const u32 j = k - d.first;
    if (j == d.n - 1) { const ag_crun g = r[d.nruns - 1]; const u32 tp = g.dst + g.len - 1; chain_pos[k] = tp; chain_base[k] = ref[tp]; return; }
    const u32 b = d.F + j;
    u32 a = 0, e = d.nruns;   // last run with src <= b
    while (e - a > 1) { const u32 mid = (a + e) >> 1; if (r[mid].src <= b) a = mid; else e = mid; }
    const ag_crun g = r[a];
    chain_pos[k] = b < g.src + g.len ? g.dst + (b - g.src) : n_ref + g.gap_tail + (b - g.src - g.len);
    char c = d.fr ? blob[d.base_off + (d.size - 1 - b)] : blob[d.base_off + b];
    if (d.fr) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    chain_base[k] = (unsigned char)c;
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_cm1: per-position summary of the contiMer table (one 8-byte load per lookup in the sweeps)
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void k_cm1(DevView d, ag_cm1* out, unsigned char* pos_term, u32* many, u32* brk) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.n_pos) return;
    const ag_cm1 c = ag_make_cm1(d.cmt, p);
    out[p] = c; many[p] = c.cid == AG_CM_MANY ? 1u : 0u;
    brk[p] = p ? ag_cm1_break(ag_make_cm1(d.cmt, p - 1), c) : 0u;
    unsigned char t = 0;
    for (u32 e = d.cmt.start[p]; e < d.cmt.start[p + 1]; e++) if (d.cmt.cm[e].chain == d.cmt.cm[e].term) t = 1;
    pos_term[p] = t;
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_build: the node sweep, with the edges of the common case settled in the same pass.
//   * one CTA per tile, 8 warps x (31 owned positions + 1 halo lane), one thread per position;
//   * the tile's alignments arrive sorted by global alignment order and are staged through shared memory in chunks (touch
//     arithmetic pre-reduced to ag_fast by k_prep, the left mates' packed bases next to them); each warp keeps only the entries that
//     overlap its 32 positions (ballot) and every thread replays the reference's first-compatible clustering on a node list held in
//     shared memory ([field][slot][thread], conflict-free);
//   * the item a touch resolves to never moves (lists grow at the tail), and the call that starts at q continues at q + 1 with the item
//     the NEXT lane resolves for the same alignment in the same iteration: one shuffle yields the edge (AG:1590-1623), kept as a bit per
//     successor item until the final node indices exist (k_succ).  Calls that cannot be settled this way (multi-segment CIGARs, several
//     contiMers on a side, item >= 32) flag the tile for the generic edge sweep k_edges;
//   * a tile's nodes go out as ONE block of final-format records at an atomically reserved offset (no ordering between tiles, nobody
//     waits); k_succ moves the blocks into position order while it expands the successor bits.
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef AG_NCHUNK
#define AG_NCHUNK 128
#endif
#ifndef AG_NODES_MINB
#define AG_NODES_MINB 6
#endif
#ifndef AG_EDGES_MINB
#define AG_EDGES_MINB 5
#endif
#ifndef AG_NCHUNK_NODES
#define AG_NCHUNK_NODES 128
#endif
#ifndef AG_CODE4
#define AG_CODE4 0   // 1: stage the left mates as oriented 4-bit codes (ag_code4_word) instead of their raw packed words — NOT yet measured on the GPU
#endif
constexpr int NCHUNK = AG_NCHUNK;            // chunk of tile alignments staged per round in k_edges
constexpr int NCHUNK_N = AG_NCHUNK_NODES;    // ... in k_build
constexpr int NODE_SCAP = AG_NODE_SCAP;                         // nodes per position kept in shared memory; more spill to the pool
constexpr int NODES_SMEM = AG_NF * NODE_SCAP * AG_TILE * 4;     // bytes of dynamic shared memory for the node lists
constexpr int READ_WORDS_MAX = 24;                              // stage reads of up to 256 bases (16 + 8 words) per chunk entry

__device__ __forceinline__ void emit_node(const DevView& d, u32 v, u32 q, char refb, u32 cid, u32 coff, u32 cid0, u32 coff0, u32 moff, u32 cov, const u32* cnt,
                                          u32 sread, u32 sl, u32 succ) {
    ag_nodec c; c.cid = cid; c.coff = coff; c.cid0 = cid0; c.coff0 = coff0;
    d.pool_c[v] = c;
    ag_nodew w; w.succ0 = succ; w.succ1 = AG_NONE; w.moff = moff; w.misc = ag_node_misc(cid, coff, cov, cnt, refb, d.coverage);  // succ0 holds the item mask until k_succ
    d.pool_w[v] = w;
    *reinterpret_cast<uint2*>(d.pool_sref + 2 * (size_t)v) = make_uint2(sread, sl);
    d.pool_pos[v] = q;
    if (d.pool_cc) { u32* o = d.pool_cc + 6 * (size_t)v; o[0] = cov; for (int j = 0; j < 5; j++) o[1 + j] = cnt[j]; }
}

__global__ void __launch_bounds__(AG_TILE, AG_NODES_MINB) k_build(DevView d) {
    extern __shared__ u32 s_dyn[];                                  // [AG_NF][NODE_SCAP][AG_TILE] node slots, then [NCHUNK_N][rw] staged read words
    __shared__ ag_fast s_f[NCHUNK_N];
    __shared__ u32 s_idx[NCHUNK_N];
    __shared__ u32 s_scan[33];
    __shared__ u32 s_base, s_flag;
    constexpr u32 READS0 = AG_NF * NODE_SCAP * AG_TILE;             // word offset of the staged reads inside s_dyn
    if (*(volatile int*)d.err & E_KEY_CAP) return;                 // incomplete key list: the step is repeated with a larger key buffer
    const u32 tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 wq0 = tile * AG_TPOS + warp * AG_WPOS, q = wq0 + lane;   // lane 31 = halo: first position of the next warp / tile
    const bool active = q < d.n_ref, owner = active && lane < AG_WPOS;
    if (threadIdx.x == 0) s_flag = 0;
    ag_slots sv; sv.saddr = (u32)__cvta_generic_to_shared(s_dyn + threadIdx.x);
    ag_plist pl; pl.n = 0; pl.ovf_head = pl.ovf_tail = AG_NONE;
    ag_cm1 ca; ca.cid = ca.coff = AG_NONE;
    if (active) ca = d.cm1[q];
    const u32 rw = d.rw, s2 = d.reads.stride2;
    const u32 kb = d.tile_start[tile], ke = d.tile_start[tile + 1];
    for (u32 c0 = kb; c0 < ke; c0 += NCHUNK_N) {
        const u32 cn = min((u32)NCHUNK_N, ke - c0);
        if (c0 != kb) __syncthreads();                             // everybody is done with the previous chunk
#if AG_CODE4
        if (threadIdx.x < cn) { const u32 idx = d.vals[c0 + threadIdx.x]; s_idx[threadIdx.x] = idx; s_f[threadIdx.x] = d.fast[idx]; }
        if (rw) {   // oriented 4-bit codes, one (entry, word) task at a time over the whole CTA; the tasks re-read idx / read / length (L1) so one barrier suffices
            const u32 sm_ = d.reads.stridem;
            for (u32 task = threadIdx.x; task < cn * rw; task += AG_TILE) {
                const u32 e = task / rw, j = task - e * rw;
                const ag_fast* gf = d.fast + d.vals[c0 + e];
                const u32 read_rc = gf->read, len = gf->lsrc_len >> 16, read = read_rc >> 1;
                s_dyn[READS0 + task] = ag_code4_word(d.reads.bases + (u64)read * s2, d.reads.nmask + (u64)read * sm_, s2, sm_, read_rc & 1, len, j);
            }
        }
#else
        if (threadIdx.x < cn) {   // stage entry: prepared record + the left mate's packed bases and non-ACGT plane (one barrier per chunk)
            const u32 idx = d.vals[c0 + threadIdx.x];
            const ag_fast f = d.fast[idx];
            s_idx[threadIdx.x] = idx;
            s_f[threadIdx.x] = f;
            if (rw) {
                const u32 read = f.read >> 1;
                const u32* __restrict__ gb = d.reads.bases + (u64)read * s2;
                const u32* __restrict__ gm = d.reads.nmask + (u64)read * d.reads.stridem;
                u32* o = s_dyn + READS0 + threadIdx.x * rw;
                for (u32 j = 0; j < s2; j++) o[j] = gb[j];
                for (u32 j = s2; j < rw; j++) o[j] = gm[j - s2];
            }
        }
#endif
        __syncthreads();
        for (u32 r0 = 0; r0 < cn; r0 += 32) {
            // which of these 32 entries touch any of the warp's 32 positions?
            bool ov = false;
            if (r0 + lane < cn) { const u32 lo = s_f[r0 + lane].lo; ov = lo <= wq0 + 31 && lo + s_f[r0 + lane].span >= wq0; }
            u32 mask = __ballot_sync(0xFFFFFFFFu, ov);
            while (mask) {
                const u32 a = r0 + (u32)__ffs((int)mask) - 1;
                mask &= mask - 1;
                const ag_fast f = s_f[a];
                u32 item = AG_NONE; bool want = false;
                if (active && q - f.lo <= f.span) {
                    const u32 roff = READS0 + a * rw;
                    const u32 fread = f.read, flen = f.lsrc_len >> 16;
                    const ag_reads reads = d.reads;
                    auto codef = [=](u32 soff) -> int {
                        if (!rw) return reads.code(fread, flen, soff);
#if AG_CODE4
                        return (int)((s_dyn[roff + (soff >> 3)] >> ((soff & 7) * 4)) & 7u);
#endif
                        const u32 rc = fread & 1, i = rc ? flen - 1 - soff : soff;
                        if ((s_dyn[roff + s2 + (i >> 5)] >> (i & 31)) & 1) return 4;
                        const u32 c = (s_dyn[roff + (i >> 4)] >> ((i & 15) * 2)) & 3;
                        return rc ? 3 - (int)c : (int)c;
                    };
                    item = ag_lane_touch(want, pl, sv, d.ovf, d.cmt, d.cm1, ca, f, d.alnp + s_idx[a], d.ext, q, (u32)d.k, d.iv, false, codef);
                }
                // the call that starts at q continues on the item the next lane resolved for this alignment
                const u32 nb = __shfl_down_sync(0xFFFFFFFFu, item, 1);
                if (want && lane < AG_WPOS) {
                    if (item != AG_NONE && nb < 32u) ag_note_succ(pl, sv, d.ovf, item, nb);
                    else atomicOr(&s_flag, item == AG_NONE ? 1u : 2u);   // 1: not a clean alignment; 2: successor item does not fit the mask
                }
            }
        }
    }
    // ---- the tile's nodes go out as one block at an atomically reserved offset ----
    u32 total; const u32 ex = block_excl_scan(owner ? pl.n : 0u, s_scan, total);
    if (threadIdx.x == 0) {
        u32 b = total ? atomicAdd(d.pool_count, total) : 0u;
        d.tile_base[tile] = b; d.tile_nodes[tile] = total;
        if (s_flag) d.tile_flag[tile] = s_flag;
        if (total && (unsigned long long)b + total > d.node_cap) { atomicOr(d.err, E_NODE_CAP); b = AG_NONE; }
        s_base = b;
    }
    __syncthreads();
    if (!owner) return;
    if (s_base == AG_NONE) { d.pos_pool[q] = 0; return; }
    u32 v = s_base + ex;
    d.pos_pool[q] = v;
    if (!pl.n) return;
    const char refb = (char)d.ref[q];
    if (ca.cid != AG_CM_MANY) {
        const u32 nloc = pl.n < (u32)NODE_SCAP ? pl.n : (u32)NODE_SCAP;
        for (u32 i = 0; i < nloc; i++, v++) {
            u32 cnt[5];
            for (u32 j = 0; j < 5; j++) cnt[j] = sv.ld(AG_F_CNT + j, i);
            emit_node(d, v, q, refb, ca.cid, ca.coff, sv.ld(AG_F_CID0, i), sv.ld(AG_F_COFF0, i), sv.ld(AG_F_MOFF, i), sv.ld(AG_F_COV, i), cnt, sv.ld(AG_F_SREAD, i),
                      sv.ld(AG_F_SL, i), sv.ld(AG_F_SUCC, i));
        }
    }
    for (u32 o = pl.ovf_head; o != AG_NONE; o = d.ovf.next[o], v++) {
        const ag_nodeb b = d.ovf.node[o];
        emit_node(d, v, q, refb, b.cid, b.coff, b.cid0, b.coff0, b.moff, b.cov, b.cnt, b.sread, b.soff_len, b.succ);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_stage + k_build_tma: the node sweep fed by bulk-asynchronous (TMA) copies.
//   k_stage gathers, per tile key in sorted order, everything the sweep needs about that alignment into ONE contiguous record: the 48-byte
//   prepared header (ag_fast) followed by the left mate as ORIENTED 4-bit codes, eight per word (ag_code4_word).  A tile's records are then
//   a single contiguous, 16-byte aligned byte range, which k_build_tma brings into shared memory with cp.async.bulk (1-D TMA) signalled
//   through an mbarrier — two chunks in flight — instead of every thread chasing vals -> fast -> reads with dependent scalar loads.  The
//   base of a touch is one shift + mask on a staged word; the mate-side match fields of a LINEAR alignment are register arithmetic.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int ST_CH = 48;                    // records per staged chunk (two chunks resident)
constexpr int ST_HDR_W = 12;                 // header words (sizeof(ag_fast) / 4)
static_assert(sizeof(ag_fast) == ST_HDR_W * 4, "staged header = ag_fast");

__global__ void k_stage(DevView d, u32* __restrict__ stage, u32 rsw) {
    if (*(volatile int*)d.err & E_KEY_CAP) return;
    const ag_reads rd = d.reads;
    const u64 nk = min(*d.nk_ptr, d.key_cap);
    const u32 vecs = rsw >> 2;                                      // 16-byte vectors per record
    // eight lanes per key, one 16-byte vector each (more for long reads): a record goes out as one contiguous 16 x vecs byte burst
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x, stride = (u64)gridDim.x * blockDim.x; g < nk * 8; g += stride) {
        const u64 key = g >> 3;
        const u32 idx = d.vals[key];
        const ag_fast* f = d.fast + idx;
        uint4* dst = reinterpret_cast<uint4*>(stage + key * rsw);
        for (u32 part = (u32)(g & 7); part < vecs; part += 8) {
            if (part < ST_HDR_W / 4) { dst[part] = reinterpret_cast<const uint4*>(f)[part]; continue; }
            const u32 read_rc = f->read, len = f->lsrc_len >> 16, read = read_rc >> 1, rc = read_rc & 1, nw = (len + 7) >> 3;
            const u32* b = rd.bases + (u64)read * rd.stride2; const u32* mk = rd.nmask + (u64)read * rd.stridem;
            const u32 j = (part - ST_HDR_W / 4) * 4;
            uint4 w;
            w.x = j + 0 < nw ? ag_code4_word(b, mk, rd.stride2, rd.stridem, rc, len, j + 0) : 0u;
            w.y = j + 1 < nw ? ag_code4_word(b, mk, rd.stride2, rd.stridem, rc, len, j + 1) : 0u;
            w.z = j + 2 < nw ? ag_code4_word(b, mk, rd.stride2, rd.stridem, rc, len, j + 2) : 0u;
            w.w = j + 3 < nw ? ag_code4_word(b, mk, rd.stride2, rd.stridem, rc, len, j + 3) : 0u;
            dst[part] = w;
        }
    }
}

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst_smem, const void* src, u32 bytes, u32 bar) {   // 1-D TMA: global -> shared, completion counted in bytes on the mbarrier
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory accesses of the sweep's inner loop through 32-bit shared-space addresses (no generic-pointer arithmetic, no cvta in the loop)
__device__ __forceinline__ u32 lds1(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds2(u32 a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds4(u32 a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts1(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
constexpr u32 SLOT_FB = AG_NODE_SCAP * AG_TILE * 4;   // bytes between two fields of a thread's slots
constexpr u32 SLOT_NB = AG_TILE * 4;                  // bytes between its two slots
// |a - b| <= thr as the reference computes it on unsigned operands (abs((int)(a - b)), AG:1296): one add and one unsigned compare (thr < 2^30)
__device__ __forceinline__ bool within(u32 a, u32 b, u32 thr) { return (a - b + thr) <= 2u * thr; }
// clauses 2 and 3 of compatible() (AG:1300-1310) against slot `s` of this thread
__device__ __forceinline__ bool slot_compatible(u32 saddr, u32 s, u32 cid0, u32 coff0, u32 moff, u32 thr) {
    const u32 b = saddr + s * SLOT_NB;
    const u32 ycid0 = lds1(b + AG_F_CID0 * SLOT_FB), ycoff0 = lds1(b + AG_F_COFF0 * SLOT_FB), ymoff = lds1(b + AG_F_MOFF * SLOT_FB);
    const bool c2 = cid0 == AG_NONE || ycid0 == AG_NONE || cid0 != ycid0 || within(coff0, ycoff0, thr);
    const bool c3 = moff == AG_NONE || ymoff == AG_NONE || within(moff, ymoff, thr);
    return c2 && c3;
}
__device__ __forceinline__ void slot_bump(u32 saddr, u32 s, u32 code) {   // coverage and the base counter of an ordinary (k1) touch
    const u32 b = saddr + s * SLOT_NB;
    sts1(b + AG_F_COV * SLOT_FB, lds1(b + AG_F_COV * SLOT_FB) + 1u);
    const u32 c = b + (AG_F_CNT + code) * SLOT_FB;
    sts1(c, lds1(c) + 1u);
}

#ifndef AG_TMA_MINB
#define AG_TMA_MINB 5
#endif
__global__ void __launch_bounds__(AG_TILE, AG_TMA_MINB) k_build_tma(DevView d, const u32* __restrict__ stage, u32 rsw) {
    extern __shared__ __align__(128) u32 s_dyn[];                   // [AG_NF][NODE_SCAP][AG_TILE] node slots, then 2 x [ST_CH][rsw] staged records
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ u32 s_scan[33];
    __shared__ u32 s_base, s_flag;
    if (*(volatile int*)d.err & E_KEY_CAP) return;
    constexpr u32 STAGE0 = AG_NF * NODE_SCAP * AG_TILE;             // word offset of the staged chunks inside s_dyn (a multiple of 4 words)
    const u32 tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 wq0 = tile * AG_TPOS + warp * AG_WPOS, q = wq0 + lane;   // lane 31 = halo: first position of the next warp / tile
    const bool active = q < d.n_ref, owner = active && lane < AG_WPOS;
    const u32 kb = d.tile_start[tile], ke = d.tile_start[tile + 1];
    const u32 nchunks = (ke - kb + ST_CH - 1) / ST_CH;
    const u32 bar0 = (u32)__cvta_generic_to_shared(&s_bar[0]), stage_s = (u32)__cvta_generic_to_shared(s_dyn + STAGE0);
    const u32 chunk_bytes = ST_CH * rsw * 4;
    auto issue = [&](u32 c) {   // one elected thread: arm the chunk's barrier with its byte count, start the bulk copy
        const u32 n = min((u32)ST_CH, ke - (kb + c * ST_CH)), bytes = n * rsw * 4, bar = bar0 + 8 * (c & 1);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stage_s + (c & 1) * chunk_bytes, stage + (size_t)(kb + c * ST_CH) * rsw, bytes, bar);
    };
    if (threadIdx.x == 0) {
        s_flag = 0;
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) { if (nchunks > 0) issue(0); if (nchunks > 1) issue(1); }
    ag_slots sv; sv.saddr = (u32)__cvta_generic_to_shared(s_dyn + threadIdx.x);
    ag_plist pl; pl.n = 0; pl.ovf_head = pl.ovf_tail = AG_NONE;
    ag_cm1 ca; ca.cid = ca.coff = AG_NONE;
    if (active) ca = d.cm1[q];
    const u32 kmer = (u32)d.k;
    // values the loop keeps in registers (opaque to the compiler, which would otherwise re-derive them from special registers / constants every iteration)
    u32 qp = active ? q : 0xFFFFFFF0u, wq = wq0, saddr = sv.saddr, edge_lane = lane < AG_WPOS ? 1u : 0u, thr = (u32)(2 * d.iv + 5 * AG_EP), rs_bytes = rsw * 4, lane_p = lane;
    asm volatile("" : "+r"(qp), "+r"(wq), "+r"(saddr), "+r"(edge_lane), "+r"(thr), "+r"(rs_bytes), "+r"(lane_p));
    const bool slots_mode = ca.cid != AG_CM_MANY;
    for (u32 c = 0; c < nchunks; c++) {
        const u32 cn = min((u32)ST_CH, ke - (kb + c * ST_CH));
        mbar_wait(bar0 + 8 * (c & 1), (c >> 1) & 1);
        u32 buf_s = stage_s + (c & 1) * chunk_bytes;                  // shared-space address of this chunk's records
        asm volatile("" : "+r"(buf_s));
        for (u32 r0 = 0; r0 < cn; r0 += 32) {
            // which of these 32 records touch any of the warp's 32 positions?
            bool ov = false;
            if (r0 + lane_p < cn) { const uint2 ls = lds2(buf_s + (r0 + lane_p) * rs_bytes); ov = ls.x <= wq + 31 && ls.x + ls.y >= wq; }
            u32 mask = __ballot_sync(0xFFFFFFFFu, ov);
            while (mask) {
                const u32 a = r0 + (u32)__ffs((int)mask) - 1;
                mask &= mask - 1;
                const u32 rec = buf_s + a * rs_bytes;
                const uint2 ls = lds2(rec);                          // lo, span
                const u32 dq = qp - ls.x;
                u32 item = AG_NONE; bool want = false;
                if (dq <= ls.y) {                                   // (inactive lanes carry a position no alignment reaches)
                    const uint4 h1 = lds4(rec + 16);                // mlen, mdelta, read, flags
                    if (slots_mode && (h1.w & (AG_FAST_CLEAN | AG_FAST_LINEAR)) == (AG_FAST_CLEAN | AG_FAST_LINEAR)) {
                        // the common case: one candidate, its mate-side fields by arithmetic; first-compatible scan over the two shared-memory slots
                        const uint2 h0 = lds2(rec + 8);             // lsrc | len << 16, mlo
                        const uint2 h2 = lds2(rec + 32);            // mcid0, mcd
                        const bool pm = (qp - h0.y) < h1.x;
                        const u32 moff = pm ? qp + h1.y : AG_NONE;
                        const u32 cid0 = pm ? h2.x : AG_NONE;
                        const u32 coff0 = cid0 != AG_NONE ? qp + h2.y : AG_NONE;
                        const bool bump = dq < ls.y;                // a call starts here (kind 1); else the stand-alone k2 of the last call
                        const u32 aoff = (h0.x & 0xFFFFu) + dq;
                        u32 code = 0;
                        if (bump) code = (lds1(rec + ST_HDR_W * 4 + ((aoff >> 1) & 0xFFFCu)) >> ((aoff << 2) & 28u)) & 7u;
                        want = bump;
                        if (pl.n >= 1 && slot_compatible(saddr, 0, cid0, coff0, moff, thr)) { if (bump) slot_bump(saddr, 0, code); item = 0; }
                        else if (pl.n >= 2 && slot_compatible(saddr, 1, cid0, coff0, moff, thr)) { if (bump) slot_bump(saddr, 1, code); item = 1; }
                        else {   // a new node, or one of the (rare) nodes beyond the two slots: the general routine (it rescans the slots, harmlessly)
                            ag_nodem cm; cm.cid = ca.cid; cm.coff = ca.coff; cm.cid0 = cid0; cm.coff0 = coff0; cm.moff = moff;
                            const u32 len = h0.x >> 16, slen = bump ? kmer : ag_min_u32(kmer, len - aoff);
                            item = ag_touch_slots(pl, sv, d.ovf, cm, bump, bump && slen ? (int)code : -1, h1.z, aoff | (slen << 16), d.iv);
                        }
                    } else {
                        ag_fast f;
                        { const uint4 g0 = lds4(rec), g2 = lds4(rec + 32);
                          f.lo = g0.x; f.span = g0.y; f.lsrc_len = g0.z; f.mlo = g0.w; f.mlen = h1.x; f.mdelta = h1.y; f.read = h1.z; f.simple = h1.w; f.mcid0 = g2.x; f.mcd = g2.y; f.aln = g2.z; f.pad = 0; }
                        const u32 cws = rec + ST_HDR_W * 4;
                        auto codef = [=](u32 soff) -> int { return (int)((lds1(cws + ((soff >> 3) << 2)) >> ((soff & 7u) << 2)) & 7u); };
                        item = ag_lane_touch(want, pl, sv, d.ovf, d.cmt, d.cm1, ca, f, d.alnp + f.aln, d.ext, qp, kmer, d.iv, false, codef);
                    }
                }
                // the call that starts at q continues on the item the next lane resolved for this alignment
                const u32 nb = __shfl_down_sync(0xFFFFFFFFu, item, 1);
                if (want && edge_lane) {
                    if (item < (u32)AG_NODE_SCAP && nb < 32u) { const u32 sa = saddr + item * SLOT_NB + AG_F_SUCC * SLOT_FB; sts1(sa, lds1(sa) | (1u << nb)); }
                    else if (item != AG_NONE && nb < 32u) ag_note_succ(pl, sv, d.ovf, item, nb);
                    else atomicOr(&s_flag, item == AG_NONE ? 1u : 2u);   // 1: not a clean alignment; 2: successor item does not fit the mask
                }
            }
        }
        if (c + 2 < nchunks) {   // everybody is done with this buffer: refill it with the chunk after next
            __syncthreads();
            if (threadIdx.x == 0) issue(c + 2);
        }
    }
    // ---- the tile's nodes go out as one block at an atomically reserved offset ----
    u32 total; const u32 ex = block_excl_scan(owner ? pl.n : 0u, s_scan, total);
    if (threadIdx.x == 0) {
        u32 b = total ? atomicAdd(d.pool_count, total) : 0u;
        d.tile_base[tile] = b; d.tile_nodes[tile] = total;
        if (s_flag) d.tile_flag[tile] = s_flag;
        if (total && (unsigned long long)b + total > d.node_cap) { atomicOr(d.err, E_NODE_CAP); b = AG_NONE; }
        s_base = b;
    }
    __syncthreads();
    if (!owner) return;
    if (s_base == AG_NONE) { d.pos_pool[q] = 0; return; }
    u32 v = s_base + ex;
    d.pos_pool[q] = v;
    if (!pl.n) return;
    const char refb = (char)d.ref[q];
    if (ca.cid != AG_CM_MANY) {
        const u32 nloc = pl.n < (u32)NODE_SCAP ? pl.n : (u32)NODE_SCAP;
        for (u32 i = 0; i < nloc; i++, v++) {
            u32 cnt[5];
            for (u32 j = 0; j < 5; j++) cnt[j] = sv.ld(AG_F_CNT + j, i);
            emit_node(d, v, q, refb, ca.cid, ca.coff, sv.ld(AG_F_CID0, i), sv.ld(AG_F_COFF0, i), sv.ld(AG_F_MOFF, i), sv.ld(AG_F_COV, i), cnt, sv.ld(AG_F_SREAD, i),
                      sv.ld(AG_F_SL, i), sv.ld(AG_F_SUCC, i));
        }
    }
    for (u32 o = pl.ovf_head; o != AG_NONE; o = d.ovf.next[o], v++) {
        const ag_nodeb b = d.ovf.node[o];
        emit_node(d, v, q, refb, b.cid, b.coff, b.cid0, b.coff0, b.moff, b.cov, b.cnt, b.sread, b.soff_len, b.succ);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_posfix / k_succ: tile blocks -> position order (final index = block's rank offset + index inside the block), successor-item
// bits -> successor node indices, subject to the contig-consistency predicate of AG:1600-1615
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void k_posfix(DevView d) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > d.n_pos) return;
    AG_BAIL(d);
    const u32 n_nodes = min(*d.nn_ptr, d.node_cap);
    if (q >= d.n_ref) { d.pos_node[q] = n_nodes; return; }   // the contig-insertion positions behind the unit hold no nodes
    const u32 t = q / AG_TPOS;
    d.pos_node[q] = d.pos_pool[q] - d.tile_base[t] + d.tile_prefix[t];
}
__global__ void k_succ(DevView d) {
    AG_BAIL(d);
    AG_FOR_N(v, min(*d.nn_ptr, d.node_cap)) {   // v: index in sweep order
        const u32 q = d.pool_pos[v], t = q / AG_TPOS;
        const u32 f = v - d.tile_base[t] + d.tile_prefix[t];   // index in position order
        ag_nodew w = d.pool_w[v];
        const ag_nodec x = d.pool_c[v];
        u32 mask = w.succ0, head = AG_NONE;
        w.succ0 = AG_NONE;
        if (mask) {
            const u32 p1 = d.pos_pool[q + 1], f1 = d.pos_node[q + 1];
            while (mask) {
                const u32 j = (u32)__ffs((int)mask) - 1;
                mask &= mask - 1;
                if (!ag_edge_ok_c(x, d.pool_c[p1 + j], d.iv)) continue;
                const u32 tg = f1 + j;
                if (w.succ0 == AG_NONE) w.succ0 = tg;
                else if (w.succ1 == AG_NONE) w.succ1 = tg;
                else {
                    const u32 o = atomicAdd(d.eovf_count, 1u);
                    if (o >= d.eovf_cap) { atomicOr(d.err, E_EDGE_OVF); break; }
                    d.eovf_target[o] = tg; d.eovf_next[o] = head; head = o;
                    w.misc |= AG_NW_OVF;
                }
            }
        }
        if (head != AG_NONE) d.eovf_head[f] = head;
        d.node_c[f] = x; d.node_w[f] = w; d.node_pos[f] = q;
        *reinterpret_cast<uint2*>(d.node_sref + 2 * (size_t)f) = *reinterpret_cast<const uint2*>(d.pool_sref + 2 * (size_t)v);
        if (d.node_cc) for (int j = 0; j < 6; j++) d.node_cc[6 * (size_t)f + j] = d.pool_cc[6 * (size_t)v + j];
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_edges: generic edge sweep over the tiles k_build flagged, against the FINAL table: for every call starting at q resolve the
// candidates on both sides to their first-compatible nodes and add the edge once  (AG:1590-1623)
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_edge(const DevView& d, u32 v, u32 tgt) {
    ag_nodew* w = &d.node_w[v];
    if (w->succ0 == tgt || w->succ1 == tgt) return;
    if (w->succ0 == AG_NONE) { w->succ0 = tgt; return; }
    if (w->succ1 == AG_NONE) { w->succ1 = tgt; return; }
    if (w->misc & AG_NW_OVF) { for (u32 o = d.eovf_head[v]; o != AG_NONE; o = d.eovf_next[o]) if (d.eovf_target[o] == tgt) return; }
    u32 o = atomicAdd(d.eovf_count, 1u);
    if (o >= d.eovf_cap) { atomicOr(d.err, E_EDGE_OVF); return; }
    d.eovf_target[o] = tgt;
    d.eovf_next[o] = (w->misc & AG_NW_OVF) ? d.eovf_head[v] : AG_NONE;
    d.eovf_head[v] = o;
    w->misc |= AG_NW_OVF;
}

// one candidate when both sides have at most one contiMer (the common case), else the reference's nested enumeration
template <class F> __device__ __forceinline__ void for_candidates_fast(const DevView& d, u32 q, const ag_cm1& ca, u32 mate, F f) {
    ag_cm1 cb; cb.cid = cb.coff = AG_NONE;
    if (mate != AG_NONE) cb = d.cm1[mate];
    if (ca.cid != AG_CM_MANY && cb.cid != AG_CM_MANY) {
        ag_nodem c; c.cid = ca.cid; c.coff = ca.coff; c.cid0 = cb.cid; c.coff0 = cb.coff; c.moff = mate;
        f(c);
    } else ag_for_candidates(d.cmt, q, mate, f);
}

__global__ void __launch_bounds__(AG_TILE, AG_EDGES_MINB) k_edges(DevView d) {
    __shared__ ag_fast s_f[NCHUNK];
    __shared__ u32 s_idx[NCHUNK];
    const u32 tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tflag = d.tile_flag[tile];
    if (!tflag) return;
    AG_BAIL(d);
    const bool all = (tflag & 2u) != 0;                   // else: only the alignments that are not clean
    const u32 wq0 = tile * AG_TPOS + warp * AG_WPOS, q = wq0 + lane;
    u32 nb0 = 0, nn0 = 0;
    if (q < d.n_ref && lane < AG_WPOS) { nb0 = d.pos_node[q]; nn0 = d.pos_node[q + 1] - nb0; }
    const bool active = nn0 != 0;
    ag_cm1 ca; ca.cid = ca.coff = AG_NONE;
    ag_cm1 ca1 = ca;
    if (active) { ca = d.cm1[q]; ca1 = d.cm1[q + 1]; }
    u32 last_v = AG_NONE, last_t = AG_NONE;               // the previous edge: most touches repeat it
    const u32 kb = d.tile_start[tile], ke = d.tile_start[tile + 1];
    for (u32 c0 = kb; c0 < ke; c0 += NCHUNK) {
        const u32 cn = min((u32)NCHUNK, ke - c0);
        if (threadIdx.x < cn) {
            const u32 idx = d.vals[c0 + threadIdx.x];
            s_idx[threadIdx.x] = idx;
            s_f[threadIdx.x] = d.fast[idx];
        }
        __syncthreads();
        for (u32 r0 = 0; r0 < cn; r0 += 32) {
            bool ov = false;
            if (r0 + lane < cn) { const u32 lo = s_f[r0 + lane].lo; ov = lo <= wq0 + 30 && lo + s_f[r0 + lane].span >= wq0 && (all || !(s_f[r0 + lane].simple & AG_FAST_CLEAN)); }
            u32 mask = __ballot_sync(0xFFFFFFFFu, ov);
            while (mask) {
                const u32 a = r0 + (u32)__ffs((int)mask) - 1;
                mask &= mask - 1;
                const ag_fast f = s_f[a];
                if (!active || q - f.lo >= f.span + (f.simple ? 0u : 1u)) continue;   // simple: calls start at lo .. lo+span-1 only
                ag_touch t;
                if (f.simple) t = ag_fast_touch(f, q, (u32)d.k);
                else { t = ag_locate(d.alnp[s_idx[a]], d.ext, q, (u32)d.k); if (t.kind != 1) continue; }
                const u32 nb1 = d.pos_node[t.npos], nn1 = d.pos_node[t.npos + 1] - nb1;
                const ag_cm1 cn1 = (t.npos == q + 1) ? ca1 : d.cm1[t.npos];
                for_candidates_fast(d, q, ca, t.mate, [&](const ag_nodem& c) {
                    const u32 ci = ag_first_compatible(d.node_c, d.node_w, nb0, nn0, c, d.iv);
                    if (ci == AG_NONE) return;
                    const ag_nodec x = d.node_c[nb0 + ci];
                    for_candidates_fast(d, t.npos, cn1, t.nmate, [&](const ag_nodem& c2) {
                        const u32 ni = ag_first_compatible(d.node_c, d.node_w, nb1, nn1, c2, d.iv);
                        if (ni == AG_NONE) return;
                        if (nb0 + ci == last_v && nb1 + ni == last_t) return;
                        if (ag_edge_ok_c(x, d.node_c[nb1 + ni], d.iv)) add_edge(d, nb0 + ci, nb1 + ni);
                        last_v = nb0 + ci; last_t = nb1 + ni;
                    });
                });
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// walk: components (union-find), per-component replay, sequential fallback, materialisation
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 uf_find(u32* parent, u32 x) {
    for (;;) {
        u32 p = parent[x];
        if (p == x) return x;
        u32 gp = parent[p];
        if (gp != p) atomicCAS(&parent[x], p, gp);
        x = p;
    }
}
__device__ __forceinline__ void uf_unite(u32* parent, u32 a, u32 b) {
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { u32 t = a; a = b; b = t; }
        if (atomicCAS(&parent[b], b, a) == b) return;
    }
}

__global__ void k_uf_init(DevView d) {
    AG_BAIL(d);
    AG_FOR_N(v, *d.nn_ptr) { d.parent[v] = v; d.cmin[v] = AG_NONE; d.cmax[v] = 0; d.walk_next[v] = AG_NONE; d.fprev[v] = AG_NONE; d.indeg[v] = 0; }
}
// Components over chain TAILS.  Only a chain's tail has live successors outside its chain and only a tail can
// leave through a contiMer detour (interior nodes always see exactly one untraversed successor), so the relations a walk can follow are:
// tail -> chains of its live successors (AG:2022-2032) and tail -> chains of the live nodes at its detour's terminal position
// (AG:2093-2114).  One thread per start candidate (= chain head), ~1/66 of the nodes.
__global__ void k_uf_tails(DevView d) {   // (after k_hrec: the hop and detour records already hold the tail's successors and where its detour lands)
    AG_BAIL(d);
    AG_FOR_N(i, *d.ncand_ptr) {
    const u32 v = d.cand_node[i];
    const ag_hrec h = d.hrec[v];
    const u32 t = h.tail;
    if (h.ts0 != AG_NONE && !(d.node_w[h.ts0].misc & AG_NW_FILTERED)) uf_unite(d.parent, t, d.chain[h.ts0].tail);
    if (h.ts1 != AG_NONE && !(d.node_w[h.ts1].misc & AG_NW_FILTERED)) uf_unite(d.parent, t, d.chain[h.ts1].tail);
    if (h.tmisc & AG_NW_OVF)
        for (u32 o = d.eovf_head[t]; o != AG_NONE; o = d.eovf_next[o]) { u32 s = d.eovf_target[o]; if (!(d.node_w[s].misc & AG_NW_FILTERED)) uf_unite(d.parent, t, d.chain[s].tail); }
    if (h.tcm == AG_NONE) continue;
    const ag_hdet dt = d.hdet[v];
    for (u32 x = dt.first; x < dt.first + dt.n; x++) if (!(d.node_w[x].misc & AG_NW_FILTERED)) uf_unite(d.parent, t, d.chain[x].tail);
    }
}
__global__ void k_uf_flatten(DevView d) {
    AG_BAIL(d);
    AG_FOR_N(i, *d.ncand_ptr) {
        u32 r = uf_find(d.parent, d.chain[d.cand_node[i]].tail);
        d.cand_label[i] = r;
        atomicMin(&d.cmin[r], i); atomicMax(&d.cmax[r], i);
    }
}

__device__ __forceinline__ ag_walkctx make_ctx(const DevView& d) {
    ag_walkctx w;
    w.nw = d.node_w; w.node_pos = d.node_pos; w.pos_node = d.pos_node; w.ovf_head = d.eovf_head; w.ovf_target = d.eovf_target;
    w.ovf_next = d.eovf_next; w.cmt = d.cmt; w.chain_pos = d.chain_pos; w.walk_next = d.walk_next; w.chain = d.chain; w.hrec = d.hrec; w.hdet = d.hdet; w.msuf = d.msuf; w.mnode = d.mnode; w.fprev = d.fprev;
    return w;
}
__device__ __forceinline__ void push_walk(const DevView& d, ag_walk r) {
    u32 o = atomicAdd(d.walk_count, 1u);
    if (o >= d.walk_cap) { atomicOr(d.err, E_WALK_CAP); return; }
    r.tail_sread = d.node_sref[2 * (size_t)r.last_node]; r.tail_soff_len = d.node_sref[2 * (size_t)r.last_node + 1];
    d.walks[o] = r;
}

// ---- forced-link chains: in-degrees, links, list ranking (Wyllie pointer jumping, double-buffered) ------------------------------
__global__ void k_indeg(DevView d) {
    AG_BAIL(d);
    AG_FOR_N(v, *d.nn_ptr) {
    ag_nodew w = d.node_w[v];
    if (w.misc & AG_NW_FILTERED) continue;
    if (w.succ0 != AG_NONE && !(d.node_w[w.succ0].misc & AG_NW_FILTERED)) atomicAdd(&d.indeg[w.succ0], 1u);
    if (w.succ1 != AG_NONE && !(d.node_w[w.succ1].misc & AG_NW_FILTERED)) atomicAdd(&d.indeg[w.succ1], 1u);
    if (w.misc & AG_NW_OVF) for (u32 o = d.eovf_head[v]; o != AG_NONE; o = d.eovf_next[o]) { u32 s = d.eovf_target[o]; if (!(d.node_w[s].misc & AG_NW_FILTERED)) atomicAdd(&d.indeg[s], 1u); }
    }
}
// one global pointer-jumping round (double-buffered); `last`: a record that is still open afterwards asks for more rounds (E_RANK_MORE:
// the step is repeated with a larger round count, remembered by the context)
// (cand_flag, last round only: 1 for the live nodes that are not chain-interior = the walk's start candidates, ready for the compaction scan)
__global__ void k_rank(const ag_chain* __restrict__ in, ag_chain* __restrict__ out, const u32* __restrict__ nn_ptr, int* err, int last,
                       const ag_nodew* __restrict__ nw, u32* __restrict__ cand_flag) {
    if (*(volatile int*)err & (E_FATAL & ~E_RANK_MORE)) return;
    AG_FOR_N(v, *nn_ptr) {
        ag_chain c = in[v];
        if (c.jump != AG_NONE) {
            ag_chain j = in[c.jump];
            c.len += j.len; c.flg += j.flg; c.tail = j.tail; c.jump = j.jump;
            if (last && c.jump != AG_NONE) atomicOr(err, E_RANK_MORE);
        }
        out[v] = c;
        if (cand_flag) cand_flag[v] = (nw[v].misc & (AG_NW_FILTERED | AG_NW_INTERIOR)) ? 0u : 1u;
    }
}

// forced links + list ranking, step 1, in one kernel: a block of 1024 consecutive nodes derives its nodes' forced links (k_links' work: fnext,
// fprev, the INTERIOR mark) and contracts the links that stay inside the block by pointer jumping in shared memory.  Forced links point to higher
// node indices and chains are short-range (the next node is the next position), so almost every chain is finished here; what remains are links
// that leave the block, resolved by a few global k_rank rounds.
// The rounds are bound by shared-memory traffic, so the block-local record is ONE 64-bit word {jump: 11 bits (local index, 0x7FF = closed), tail: 10,
// len: 11, flg: 11}; it is updated in place (read phase, barrier, write phase) and only by threads whose record is still open — the open
// fraction of a chain of length L after r rounds is (L - 2^r) / L.  A link that leaves the block is kept per node (s_ext) and attached to
// the finished records through their local tail.
__global__ void __launch_bounds__(1024) k_links_rank_local(DevView d, ag_chain* recs) {
    __shared__ unsigned long long s_rec[1024];
    __shared__ u32 s_ext[1024];
    const u32 n_nodes = *d.nn_ptr;
    const u32 b0 = blockIdx.x * 1024u, t = threadIdx.x, v = b0 + t;
    if (b0 >= n_nodes || (*(volatile const int*)d.err & E_FATAL)) return;
    constexpr u32 CLOSED = 0x7FFu;
    auto pack = [](u32 jump, u32 tail, u32 len, u32 flg) { return (unsigned long long)jump | ((unsigned long long)tail << 11) | ((unsigned long long)len << 21) | ((unsigned long long)flg << 32); };
    u32 ext = AG_NONE, jl = CLOSED, len = 0, flg = 0;
    if (v < n_nodes) {
        const u32 w = ag_forced_succ(d.node_w, d.eovf_head, d.eovf_target, d.eovf_next, d.indeg, d.pos_term, d.node_pos, v);
        d.fnext[v] = w;
        if (w != AG_NONE) { atomicOr(&d.node_w[w].misc, AG_NW_INTERIOR); d.fprev[w] = v; }
        if (w != AG_NONE && w - b0 < 1024u) jl = w - b0; else ext = w;
        len = 1; flg = (d.node_w[v].misc & AG_NW_HASCONTIG) ? 1u : 0u;
    }
    unsigned long long me = pack(jl, t, len, flg);
    s_rec[t] = me; s_ext[t] = ext;
    __syncthreads();
    for (int r = 0; r < 10; r++) {   // (2^10 = block size: enough for a chain through the whole block)
        const u32 j = (u32)me & 0x7FFu;
        const bool open = j != CLOSED;
        unsigned long long o = 0;
        if (open) o = s_rec[j];
        if (!__syncthreads_or(open ? 1 : 0)) break;          // every read of this round is done (and: nobody is open any more)
        if (open) {   // me = me . o : jump and tail from o, lengths and contig flags add up (len, flg <= 1024: no carry between the fields)
            me = ((me >> 21) + (o >> 21)) << 21 | (o & 0x1FFFFFull);
            s_rec[t] = me;
        }
        __syncthreads();
    }
    if (v < n_nodes) {
        const u32 tl = (u32)(me >> 11) & 0x3FFu;
        ag_chain c; c.jump = s_ext[tl]; c.tail = b0 + tl; c.len = (u32)(me >> 21) & 0x7FFu; c.flg = (u32)(me >> 32) & 0x7FFu;
        recs[v] = c;
    }
}

// walk starts can only be live nodes that are not chain-interior: compact them (in node order) so that a component's replay does not
// have to step over every node of its range
__global__ void k_cand_flag(DevView d, u32* flag) {   // covers the whole node capacity: entries beyond the node count are 0 for the scan
    const bool dead = (*(volatile int*)d.err & E_FATAL) != 0;
    const u32 nn = dead ? 0u : *d.nn_ptr;
    AG_FOR_N(v, d.node_cap) flag[v] = (v < nn && !(d.node_w[v].misc & (AG_NW_FILTERED | AG_NW_INTERIOR))) ? 1u : 0u;
}
__global__ void k_cand_scatter(DevView d, const u32* flag) {
    AG_BAIL(d);
    if (*d.ncand_ptr > d.cand_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(d.err, E_CAND_CAP); return; }
    AG_FOR_N(v, *d.nn_ptr) if (flag[v]) d.cand_node[d.cand_rank[v]] = v;
}

// hop records of the chain heads (= start candidates)
__global__ void k_hrec(DevView d) {
    AG_BAIL(d);
    AG_FOR_N(i, *d.ncand_ptr) {
        const u32 v = d.cand_node[i];
        const ag_chain c = d.chain[v];
        const ag_hrec h = ag_make_hrec(c, d.node_w[c.tail], d.cmt, d.node_pos[c.tail]);
        d.hrec[v] = h;
        if (h.tcm != AG_NONE) d.hdet[v] = ag_make_hdet(d.cmt, d.chain_pos, d.pos_node, h.tcm);
    }
}

// One WARP per component (the warp of the candidate whose chain tail is the union-find root): the replay of the scan (AG:1972-1990)
// over the component's start candidates, in node order, is sequential and runs on lane 0; the critical path of the whole kernel is the
// largest component (a few hundred candidates, every one two or three dependent loads of never-touched lines).  So all 32 lanes first
// pull the records the replay is going to read — walk record, hop record, position, founder string of every candidate, and the walk
// records of the chain tails and of their successors — into L1/L2.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__global__ void k_walk_components(DevView d, u32* next_cand) {
    AG_BAIL(d);
    const u32 n_cand = *d.ncand_ptr;
    const u32 lane = threadIdx.x & 31;
    for (;;) {   // persistent warps (the candidate count is only known on the device) that draw candidates four at a time: a warp busy with a
        u32 c0 = 0;   // large component does not hold back the candidates behind it
        if (lane == 0) c0 = atomicAdd(next_cand, 4u);
        c0 = __shfl_sync(0xFFFFFFFFu, c0, 0);
        if (c0 >= n_cand) break;
      for (u32 i0 = c0; i0 < min(c0 + 4u, n_cand); i0++) {
        const u32 r = d.chain[d.cand_node[i0]].tail;
        if (d.parent[r] != r) continue;
        const u32 lo = d.cmin[r], hi = d.cmax[r];
        // A typical unit has a few hundred components of a few hundred candidates each (one per contig region) and thousands of tiny ones: the
        // kernel's duration is the replay of ONE large component on one lane.  Pulling a whole component into L1 up front does not survive
        // the other components replayed on the same SM, so the records are pulled one window of 32 candidates ahead of the replay.
        auto pull = [&](u32 base) {
            const u32 i = base + lane;
            if (i > hi || d.cand_label[i] != r) return;
            const u32 v = d.cand_node[i];
            prefetch_l1(&d.node_w[v]); prefetch_l1(&d.node_pos[v]);
            const ag_hrec h = d.hrec[v];
            prefetch_l1(&d.node_w[h.tail]); prefetch_l1(&d.node_pos[h.tail]); prefetch_l1(&d.walk_next[h.tail]); prefetch_l1(&d.node_sref[2 * (size_t)h.tail]);
            if (h.ts0 != AG_NONE) { prefetch_l1(&d.node_w[h.ts0]); prefetch_l1(&d.hrec[h.ts0]); }
            if (h.ts1 != AG_NONE) { prefetch_l1(&d.node_w[h.ts1]); prefetch_l1(&d.hrec[h.ts1]); }
            if (h.tcm != AG_NONE) { const ag_hdet dt = d.hdet[v]; if (dt.n) prefetch_l1(&d.node_w[dt.first]); }
        };
        const bool windowed = hi - lo >= 8;
        if (windowed) { pull(lo); __syncwarp(); }
        ag_walkctx w = make_ctx(d);
        for (u32 base = lo; base <= hi; base += 32) {
            if (windowed && base + 32 <= hi) pull(base + 32);
            __syncwarp();
            if (lane == 0) {
                const u32 end = min(hi, base + 31u);
                for (u32 i = base; i <= end; i++) {
                    if (d.cand_label[i] != r) continue;
                    const u32 v = d.cand_node[i];
                    if (d.node_w[v].misc & AG_NW_TRAV) continue;
                    d.walks[i] = ag_walk_from(w, v);   // slot = candidate index: no counter on the sequential path; k_walk_compact compacts the used slots in order
                    d.walk_used[i] = 1;
                }
            }
            __syncwarp();
        }
      }
    }
}

// component replay: the used slots (one per candidate that started a walk), compacted in candidate = scan order (candidates are in node
// order, so that IS the scan order of AG:1972-1990) straight into the page-locked host buffer; the founder string of the walk's last node
// is attached here, off the replay's sequential path.  rank = exclusive scan of walk_used over the candidate capacity.
__global__ void k_walk_compact(DevView d, const u32* __restrict__ rank) {
    AG_BAIL(d);
    if (rank[d.cand_cap] > d.hwalk_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(d.err, E_HWALK_CAP); return; }
    AG_FOR_N(i, *d.ncand_ptr) {
        if (!d.walk_used[i]) continue;
        ag_walk r = d.walks[i];
        r.tail_sread = d.node_sref[2 * (size_t)r.last_node]; r.tail_soff_len = d.node_sref[2 * (size_t)r.last_node + 1];
        d.walks_sorted[rank[i]] = r;
    }
}
// the compacted records to the page-locked host buffer, 8 bytes per lane: full 256-byte writes over PCIe
__global__ void k_walk_to_host(DevView d, const u32* __restrict__ rank) {
    AG_BAIL(d);
    const u32 nw = rank[d.cand_cap];
    if (nw > d.hwalk_cap) return;
    const uint2* __restrict__ src = reinterpret_cast<const uint2*>(d.walks_sorted); uint2* dst = reinterpret_cast<uint2*>(d.walks_host);
    AG_FOR_N(i, nw * (u32)(sizeof(ag_walk) / 8)) dst[i] = src[i];
}
// ---- fused extension path: emission filter, offsets and materialisation inputs on the device (no host round trip between walk and bases) ----
// Walk records arrive compacted in scan order, i.e. by non-decreasing start position, so `contain(previous emitted, this)` (AG:2176, AG:1897-1902)
// reduces to "this walk's end does not pass the furthest end seen so far": emitted[i] = (i == 0) || E[i] > max(E[0 .. i-1]), E = the end offset
// with the tail adjustment of AG:2164-2173 in the reference's own unsigned arithmetic.
__global__ void k_sel_prepare(DevView d, const u32* __restrict__ walk_rank, u32* __restrict__ E, u32* __restrict__ info) {
    if (blockIdx.x == 0 && threadIdx.x < 8) info[threadIdx.x] = 0;   // [0] emitted walks, [1] their bases, [2] > 100 kbp trigger (set by k_sel_flag / k_sel_fill, later in the stream)
    AG_BAIL(d);
    AG_FOR_N(i, min(walk_rank[d.cand_cap], d.hwalk_cap)) {
        const ag_walk r = d.walks_sorted[i];
        u32 eoff = r.eoff;
        if (((r.flags >> 1) & 3) != 1) eoff = eoff + (r.tail_soff_len >> 16) - 1;
        E[i] = eoff;
    }
}
__global__ void k_sel_flag(DevView d, const u32* __restrict__ walk_rank, const u32* __restrict__ E, const u32* __restrict__ M, u32* __restrict__ flag, u32* __restrict__ len, u32* trigger) {
    const bool dead = (*(volatile int*)d.err & E_FATAL) != 0;
    const u32 nw = dead ? 0u : min(walk_rank[d.cand_cap], d.hwalk_cap);
    AG_FOR_N(i, d.hwalk_cap) {   // covers the capacity: zeros behind the walk count for the scans
        u32 f = 0, l = 0;
        if (i < nw && (i == 0 || E[i] > M[i])) {
            const ag_walk r = d.walks_sorted[i];
            f = 1; l = r.len + ag_walk_tail_len(r);
            if (E[i] - r.soff > 100000u) atomicOr(trigger, 1u);   // a contig of more than 100 kbp switches the reference's 1000-position skip on (AG:2194-2202): exact sequential replay
        }
        flag[i] = f; len[i] = l;
    }
}
// compacted materialisation inputs of the emitted walks + their records for the host
__global__ void k_sel_fill(DevView d, const u32* __restrict__ walk_rank, const u32* __restrict__ flag, const u32* __restrict__ srank, const u32* __restrict__ soff,
                           u32* __restrict__ sel_start, u64* __restrict__ sel_off, u32* __restrict__ sel_tails, ag_walk* __restrict__ sel_walks, u32* __restrict__ sel_off32,
                           u32* __restrict__ info) {
    AG_BAIL(d);
    const u32 nw = min(walk_rank[d.cand_cap], d.hwalk_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = srank[d.hwalk_cap]; info[1] = soff[d.hwalk_cap]; }
    if (soff[d.hwalk_cap] > d.bases_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(d.err, E_BASES_CAP); return; }
    AG_FOR_N(i, nw) {
        if (!flag[i]) continue;
        const ag_walk r = d.walks_sorted[i];
        const u32 j = srank[i], tl = ag_walk_tail_len(r);
        sel_start[j] = r.start_node; sel_off[j] = soff[i]; sel_off32[j] = soff[i];
        sel_tails[3 * j] = r.tail_sread; sel_tails[3 * j + 1] = tl ? r.tail_soff_len : 0u; sel_tails[3 * j + 2] = r.len;
        sel_walks[j] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sel_off32[srank[d.hwalk_cap]] = soff[d.hwalk_cap];   // closing offset
}
// device buffer -> page-locked host buffer, 16 bytes per lane (count in 16-byte units from the device)
__global__ void k_to_host16(const uint4* __restrict__ src, uint4* __restrict__ dst, const u32* __restrict__ n_items, u32 item_bytes, u32 extra_items, const int* err) {
    if (*(volatile const int*)err & E_FATAL) return;
    const u64 n16 = ((u64)(*n_items + extra_items) * item_bytes + 15) / 16;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x, s = (u64)gridDim.x * blockDim.x; i < n16; i += s) dst[i] = src[i];
}

// everything the host wants to know about the step, gathered into one block for ONE device->host copy
__global__ void k_status(DevView d, const u32* __restrict__ counters, const u32* __restrict__ walk_rank, const u32* __restrict__ sel, u32* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    out[0] = (u32)(*d.err | *d.err_load); out[1] = counters[0]; out[2] = counters[1]; out[3] = counters[2]; out[4] = *d.nk_ptr;
    out[5] = d.ncand_ptr ? *d.ncand_ptr : 0u; out[6] = walk_rank ? walk_rank[d.cand_cap] : 0u;
    out[7] = sel ? sel[0] : 0u; out[8] = sel ? sel[1] : 0u; out[9] = sel ? sel[2] : 0u;   // fused extension: emitted walks, their bases, >100 kbp trigger
}

// fused extension, early hand-off: {error word, emitted walks, their bases, > 100 kbp trigger} straight into page-locked host memory, queued
// behind the copies of the emitted walk records and offsets and in front of the materialisation — the host formats records and headers
// while the bases are still being produced
__global__ void k_status_early(DevView d, const u32* __restrict__ sel, volatile u32* __restrict__ host) {
    if (blockIdx.x || threadIdx.x) return;
    host[0] = (u32)(*d.err | *d.err_load); host[1] = sel[0]; host[2] = sel[1]; host[3] = sel[2];
}

// exact sequential replay including the 1000-position skip (AG:2194-2202); used only when a >100 kbp contig was emitted
__global__ void k_walk_sequential(DevView d) {
    if (blockIdx.x || threadIdx.x) return;
    ag_walkctx w = make_ctx(d);
    u32 bso = AG_NONE, beo = AG_NONE, bei = AG_NONE;
    for (u32 cp = 0; cp < d.n_ref;) {
        for (u32 v = d.pos_node[cp]; v < d.pos_node[cp + 1]; v++) {
            if (ag_seq_trav(w, v)) continue;
            ag_walk r = ag_walk_from_seq(w, v);
            push_walk(d, r);
            u32 eoff = r.eoff;
            if (((r.flags >> 1) & 3) != 1) eoff = eoff + (d.node_sref[2 * (size_t)r.last_node + 1] >> 16) - 1;  // AG:2170
            bool contained = (bei == 0) && bso <= r.soff && beo >= eoff;     // contain(), AG:1897-1902 (ids are 0 once set)
            if (!contained) { bso = r.soff; beo = eoff; bei = 0; }
        }
        if (beo - bso > 100000u) { if (bei == 0 && cp + 1000 < beo) cp += 1000; else cp++; }
        else cp++;
    }
}

// sequential replay: thread per walk; inside a chain follow the forced links until the node where the walk left it (STOP)
__global__ void k_materialize_seq(DevView d, const u32* __restrict__ starts, const u64* __restrict__ offs, u32 n, unsigned char* out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char* o = out + offs[i];
    u32 v = starts[i];
    while (v != AG_NONE) {
        const u32 misc = d.node_w[v].misc;
        *o++ = (unsigned char)(misc & 0xFF);
        if (!(misc & AG_NW_STOP)) { v = d.fnext[v]; continue; }
        if (misc & AG_NW_DETOUR) {
            ag_cm m = d.cmt.cm[d.cmt.start[d.node_pos[v]]];
            for (u32 e = m.chain + 1; e <= m.term; e++) *o++ = d.chain_base[e];
        }
        v = d.walk_next[v];
    }
}

// materialisation in two steps when chains are valid: (1) one thread per emitted walk hops from chain head to chain head and emits
// work items, (2) one thread per chain item writes that chain's consensus bases, one warp per detour item copies the contig bases
struct MatItem { u32 a, n; u64 off; };  // detour item: a = first chain-major index, n = bases
// A chain belongs to at most one walk (it is marked as a whole), so "where does this chain's run of bases END in the output" can be kept
// per chain tail; every node then finds its own byte: end - (nodes from here to the tail).  One thread per node, no pointer chasing.
__global__ void k_mat_items(DevView d, const u32* __restrict__ starts, const u64* __restrict__ offs, u32 n, const u32* __restrict__ n_ptr, u32* tail_end, MatItem* detours, u32* counts, u32 cap) {
    if (n_ptr) { AG_BAIL(d); n = *n_ptr; }
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    u64 off = offs[i];
    u32 v = starts[i];
    while (v != AG_NONE) {
        const ag_chain c = d.chain[v];
        off += c.len;
        const u32 t = c.tail;
        tail_end[t] = (u32)off;
        if (d.node_w[t].misc & AG_NW_DETOUR) {
            ag_cm m = d.cmt.cm[d.cmt.start[d.node_pos[t]]];
            u32 o2 = atomicAdd(&counts[1], 1u);
            if (o2 < cap) { MatItem it; it.a = m.chain + 1; it.n = m.term - m.chain; it.off = off; detours[o2] = it; } else atomicOr(d.err, E_MAT_CAP);
            off += m.term - m.chain;
        }
        v = d.walk_next[t];
    }
    }
}
__global__ void k_mat_nodes(DevView d, u32 n_nodes, const u32* __restrict__ n_ptr, const u32* __restrict__ tail_end, unsigned char* out) {
    if (n_ptr) { AG_BAIL(d); n_nodes = *n_ptr; }
    for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < n_nodes; v += gridDim.x * blockDim.x) {
        const ag_chain c = d.chain[v];
        const u32 e = tail_end[c.tail];
        if (e != AG_NONE) out[e - c.len] = (unsigned char)(d.node_w[v].misc & 0xFF);
    }
}
__global__ void k_mat_detours(DevView d, const MatItem* __restrict__ items, const u32* counts, unsigned char* out) {
    // one CTA per detour item (a run of contig bases, typically thousands): byte copies, coalesced across the CTA
    for (u32 w = blockIdx.x; w < counts[1]; w += gridDim.x) {
        const MatItem it = items[w];
        for (u32 j = threadIdx.x; j < it.n; j += blockDim.x) out[it.off + j] = d.chain_base[it.a + j];
    }
}

// s[1..] of the last node of every selected walk that ended in the k-mer graph (AG:2164-2168); characters outside ACGT come out as 'N'
// here and are restored on the host from the reads' exception list
__global__ void k_mat_tails(DevView d, const u32* __restrict__ tails /* per walk: sread, soff_len, loop length */, const u64* __restrict__ offs, u32 n, const u32* __restrict__ n_ptr, unsigned char* out) {
    if (n_ptr) { AG_BAIL(d); n = *n_ptr; }
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 sread = tails[3 * i], sl = tails[3 * i + 1], len = tails[3 * i + 2];
        const u32 slen = sl >> 16, soff = sl & 0xFFFFu;
        if (slen < 2) continue;
        const u32 rlen = d.reads.len[sread >> 2];
        unsigned char* o = out + offs[i] + len;
        for (u32 j = 1; j < slen; j++) { int c = d.reads.code(sread, rlen, soff + j); *o++ = (unsigned char)"ACGTN"[c]; }
    }
}

__global__ void k_reset_marks(DevView d, u32 n_nodes) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    u32 m = d.node_w[v].misc & ~(AG_NW_TRAV | AG_NW_DETOUR | AG_NW_STOP);
    d.node_w[v].misc = (m & AG_NW_FILTERED) ? (m | AG_NW_TRAV) : m;
    d.walk_next[v] = AG_NONE; d.msuf[v] = 0; d.mnode[v] = AG_NONE;
}

__global__ void k_occupancy(DevView d, unsigned char* bits) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= d.n_pos) return;
    unsigned char x = 0;
    for (u32 j = 0; j < 8; j++) {
        u32 p = b * 8 + j;
        if (p >= d.n_pos) break;
        bool occ = d.cmt.start[p + 1] > d.cmt.start[p];
        if (p < d.n_ref && d.pos_node[p + 1] > d.pos_node[p]) occ = true;
        if (occ) x |= (unsigned char)(1u << j);
    }
    bits[b] = x;
}

struct Timer {   // section timer: CUDA events on the launching stream; the event pair is cached per host thread (sections never nest)
    cudaEvent_t a, b; cudaStream_t st;
    static cudaEvent_t* pair() { static thread_local cudaEvent_t ev[2] = {nullptr, nullptr}; static thread_local int dev = -1; int cur = -1; cudaGetDevice(&cur);
                                 if (!ev[0] || dev != cur) { cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]); dev = cur; } return ev; }
    Timer(cudaStream_t s) : st(s) { cudaEvent_t* e = pair(); a = e[0]; b = e[1]; cudaEventRecord(a, st); }
    float stop() { cudaEventRecord(b, st); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap_ms() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3; }
};

// Section times without stalling the stream: an event pair per section is only RECORDED while the step is being queued; the elapsed
// times are read after the step's single synchronisation point (collect()).
struct SectionTimes {
    static constexpr int MAX = 64;
    cudaEvent_t ev[2 * MAX]; float* dst[MAX]; int n = 0, made = 0, open = 0;
    int begin(cudaStream_t st, float* where) {   // sections may nest: every one owns its event pair from begin()
        if (n == MAX && open == 0) { cudaStreamSynchronize(st); collect(); }
        if (n == MAX) return -1;
        while (made <= n) { cudaEventCreate(&ev[2 * made]); cudaEventCreate(&ev[2 * made + 1]); made++; }
        const int id = n++;
        dst[id] = where; cudaEventRecord(ev[2 * id], st); open++;
        return id;
    }
    void end(cudaStream_t st, int id) { if (id >= 0) { cudaEventRecord(ev[2 * id + 1], st); open--; } }
    void collect() { for (int i = 0; i < n; i++) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]) == cudaSuccess) *dst[i] += ms; } n = 0; }   // after a sync that covers every recorded event
    void release() { for (int i = 0; i < 2 * made; i++) cudaEventDestroy(ev[i]); made = n = 0; }
};
struct Section { SectionTimes& t; cudaStream_t st; int id; Section(SectionTimes& t_, cudaStream_t s, float* where) : t(t_), st(s), id(t_.begin(s, where)) {} ~Section() { t.end(st, id); } };

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
struct AgDevice::Impl {
    cudaStream_t st = nullptr;
    // reads
    DBuf<u32> r_bases, r_nmask; DBuf<uint16_t> r_len; DBuf<u64> r_exc; ag_reads reads{}; u64 n_pairs = 0; bool reads_owned = false;
    // unit inputs
    DBuf<unsigned char> ref, chain_base; DBuf<u32> cm_start, chain_pos; DBuf<ag_cm> cm; DBuf<ag_cthread> cthreads; DBuf<ag_aln> aln; DBuf<ag_seg> ext;
    u32 n_ref = 0, n_pos = 0, n_cm = 0, n_aln = 0;
    // build products
    DBuf<ag_alnp> alnp; DBuf<ag_fast> fast; DBuf<u32> ntiles, key_off, keys, vals, keys2, vals2, hist, tile_cnt, tile_start, tile_flag, many, many_prefix, brk, lin_prefix; DBuf<u32> stage;
    DBuf<u32> tile_base, tile_nodes, tile_prefix, pos_pool, pool_sref, pool_pos, pool_cc; DBuf<ag_nodec> pool_c; DBuf<ag_nodew> pool_w;
    DBuf<ag_nodeb> ovf_node; DBuf<u32> ovf_next, counters; DBuf<int> err;
    DBuf<u32> pos_node;
    DBuf<ag_nodec> node_c; DBuf<ag_nodew> node_w; DBuf<u32> node_sref, node_pos, node_cc;
    DBuf<u32> eovf_head, eovf_target, eovf_next;
    DBuf<u32> walk_next, parent, cmin, cmax, walk_used; DBuf<ag_walk> walks, walks2; DBuf<ag_cm1> cm1; DBuf<MatItem> mat_detours; DBuf<u32> tail_end; u32 n_cand = 0; DBuf<unsigned char> pos_term; DBuf<u32> indeg, fnext, fprev, msuf, mnode, cand_rank, cand_node, cand_label; DBuf<ag_chain> chain_a, chain_b; DBuf<ag_hrec> hrec; DBuf<ag_hdet> hdet; DBuf<int> changed;
    DBuf<unsigned char> out_bases, occ; DBuf<u32> sel_start, sel_tails; DBuf<u64> sel_off;
    PinnedBuf h_walks, h_bases, h_occ, h_sel, h_s;   // h_s: page-locked landing zone of the scalar read-backs (a pageable destination makes every copy a synchronous staged transfer)
    // text ingestion (ag_ingest.cuh)
    DBuf<char> contig_blob; u64 blob_version = 0; DBuf<ag_cdesc> cdesc; DBuf<ag_crun> cruns;   // run-space contig threads
    DBuf<char> raw, exc_chr; DBuf<u32> nl, nl_blk, rlen, s_keep, s_next, s_aoff, s_eoff, s_lost, ing; DBuf<u64> exc_key; DBuf<ag_srec> srec; FileStager stager;
    u64 n_ext = 0; bool aln_ingested = false;
    u64 win_lo = 0, win_hi = ~0ull;   // pair-id window of the read set that is resident (everything by default)
    bool window_miss = false;          // the last ingest_sam met a read id outside the window
    // newline index of text[0, len) (device, 16-byte aligned, zero-padded to a multiple of 16).  fill == false: count (block offsets go to
    // nl_blk[blk_base ..]) and return the number of lines — one blocking read-back; fill == true: write the positions to nl[nl_base ..]
    u32 nl_index(cudaStream_t st, u64& launches, const char* text, size_t len, size_t blk_base, size_t nl_base, bool fill);
    Scanner scanner;
    u32 n_tiles = 0, n_keys = 0, n_nodes = 0;
    u32 node_cap = 0, ovf_cap = 0, eovf_cap = 0, eovf_cap_init = 0, walk_cap = 0;
    u32 key_cap = 0, cand_cap = 0, hwalk_cap = 0, n_walks = 0; int rank_rounds = 2;   // global list-ranking rounds queued per step (chains that leave their 1024-node block); more on request (E_RANK_MORE)
    u32 unit_n_ref = 0; u64 unit_n_aln = 0;   // the unit the capacities above were last derived for
    cudaEvent_t ev_early = nullptr;          // fused extension: emitted walk records + offsets are in host memory (the bases follow)
    PinnedBuf h_wrec;                        // compacted walk records, written by k_walk_compact straight into host memory
    DBuf<u32> status, walk_rank; DBuf<int> err_load; SectionTimes sections;
    bool key_cap_hooked = false; u32 node_cap_hook = 0, ovf_cap_hook = 0;   // test hooks (set_option): initial capacities instead of the size-derived ones
    bool build_queued = false, walk_queued = false, select_queued = false;   // queued on the stream, not yet checked by finish()
    // fused extension path: emission filter + materialisation queued behind the walk
    u32 bases_cap = 0, n_sel = 0, sel_bases = 0; bool sel_trigger = false;
    DBuf<u32> sel_E, sel_M, sel_flag, sel_len, sel_rank, sel_soff, sel_info, sel_off32; DBuf<ag_walk> sel_walks; PinnedBuf h_selw, h_selo;
    DevView view{};
    // counters layout: [0] n_nodes (pool_count), [1] ovf_count, [2] eovf_count, [3] walk_count, [4,5] materialise items
};

AgDevice::AgDevice(int device) : m_(new Impl), dev_(device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { delete m_; throw AgError{"no CUDA device: the AlignGraph B200 hot path has no CPU fallback"}; }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&m_->st, cudaStreamNonBlocking));
    stream_ = m_->st;
    m_->scanner.launches = &launches_;
    if (const char* e = getenv("AG_TMA")) tma_off_ = atoi(e) == 0;   // AG_TMA=0: node sweep with per-thread staging (k_build) instead of the bulk-async staged one
    m_->counters.ensure(8); m_->err.ensure(1); m_->h_s.ensure(256); m_->status.ensure(16); m_->err_load.ensure(1); CK(cudaMemset(m_->err_load.p, 0, sizeof(int)));
}
AgDevice::~AgDevice() {
    cudaSetDevice(dev_);
    // DBuf members are plain; free what we own
    Impl& m = *m_;
    if (m.reads_owned) { m.r_bases.release(); m.r_nmask.release(); m.r_len.release(); }
    m.r_exc.release();
    DBuf<unsigned char>* b8[] = {&m.ref, &m.chain_base, &m.out_bases, &m.occ};
    for (auto* b : b8) b->release();
    DBuf<u32>* b32[] = {&m.cm_start, &m.chain_pos, &m.ntiles, &m.key_off, &m.keys, &m.vals, &m.keys2, &m.vals2, &m.hist, &m.tile_cnt,
                        &m.tile_start, &m.tile_flag, &m.tile_base, &m.tile_nodes, &m.tile_prefix, &m.pos_pool, &m.pool_sref, &m.pool_pos, &m.pool_cc, &m.many, &m.many_prefix, &m.brk, &m.lin_prefix, &m.stage, &m.ovf_next, &m.counters, &m.pos_node, &m.node_sref, &m.node_pos, &m.node_cc, &m.eovf_head,
                        &m.eovf_target, &m.eovf_next, &m.walk_next, &m.parent, &m.cmin, &m.cmax, &m.walk_used, &m.sel_start, &m.sel_tails};
    for (auto* b : b32) b->release();
    for (int i = 0; i < 4; i++) m.scanner.lvl[i].release();
    m.scanner.state.release();
    m.cm.release(); m.cthreads.release(); m.aln.release(); m.ext.release(); m.alnp.release(); m.fast.release(); m.pool_c.release(); m.pool_w.release(); m.ovf_node.release(); m.err.release();
    m.node_c.release(); m.node_w.release(); m.walks.release(); m.walks2.release(); m.cm1.release(); m.mat_detours.release(); m.tail_end.release(); m.pos_term.release(); m.cand_rank.release(); m.cand_node.release(); m.cand_label.release(); m.fprev.release(); m.msuf.release(); m.mnode.release(); m.indeg.release(); m.fnext.release(); m.chain_a.release(); m.chain_b.release(); m.hrec.release(); m.hdet.release(); m.changed.release(); m.sel_off.release();
    m.contig_blob.release(); m.cdesc.release(); m.cruns.release(); m.raw.release(); m.exc_chr.release(); m.nl.release(); m.nl_blk.release(); m.rlen.release(); m.s_keep.release(); m.s_next.release(); m.s_aoff.release(); m.s_eoff.release(); m.s_lost.release(); m.ing.release(); m.exc_key.release(); m.srec.release(); m.stager.release();
    m.sel_E.release(); m.sel_M.release(); m.sel_flag.release(); m.sel_len.release(); m.sel_rank.release(); m.sel_soff.release(); m.sel_info.release(); m.sel_off32.release(); m.sel_walks.release(); m.h_selw.release(); m.h_selo.release();
    if (m.ev_early) { cudaEventDestroy(m.ev_early); m.ev_early = nullptr; } m.h_wrec.release(); m.walk_rank.release(); m.status.release(); m.err_load.release(); m.sections.release();
    m.h_walks.release(); m.h_bases.release(); m.h_occ.release(); m.h_sel.release(); m.h_s.release();
    if (ev_mat0_) { cudaEventDestroy((cudaEvent_t)ev_mat0_); cudaEventDestroy((cudaEvent_t)ev_mat1_); }
    if (st2_) { cudaStreamSynchronize((cudaStream_t)st2_); cudaStreamDestroy((cudaStream_t)st2_); cudaEventDestroy((cudaEvent_t)ev_main_); cudaEventDestroy((cudaEvent_t)ev_reads_); }
    if (m.st) cudaStreamDestroy(m.st);
    delete m_;
}

void AgDevice::set_params(int k, int iv, int coverage) {
    if (iv < 0 || iv > (1 << 28)) throw AgError{"--insertVariation out of range"};   // (the reference's 2 * iv + 25 is an int; the sweeps compare |a - b| <= 2 iv + 25 with one unsigned add)
    k_ = k; iv_ = iv; cov_ = coverage;
}
void AgDevice::set_option(const std::string& name, long value) {
    Impl& m = *m_;
    if (name == "node_cap") { m.node_cap_hook = (u32)std::max<long>(value, 1); m.unit_n_ref = 0; }
    else if (name == "ovf_cap") { m.ovf_cap_hook = (u32)std::max<long>(value, 1); m.unit_n_ref = 0; }
    else if (name == "key_cap") { m.key_cap = (u32)std::max<long>(value, 1); m.key_cap_hooked = true; }
    else if (name == "cand_cap") m.cand_cap = (u32)std::max<long>(value, 1);
    else if (name == "hwalk_cap") m.hwalk_cap = (u32)std::max<long>(value, 1);
    else if (name == "tma") tma_off_ = value == 0;
    else if (name == "bases_cap") m.bases_cap = (u32)std::max<long>(value, 1);
    else if (name == "fused_extend") fused_off_ = value == 0;
    else if (name == "scan_onepass") m_->scanner.one_pass = value != 0;
    else if (name == "rank_rounds") m.rank_rounds = (int)std::max<long>(value, 1);
    else if (name == "eovf_cap") m.eovf_cap_init = (u32)std::max<long>(value, 1);
    else if (name == "section_timing") section_timing_ = value != 0;
}

void AgDevice::pin(const void* p, size_t bytes) {
    if (!p || !bytes) return;
    CK(cudaSetDevice(dev_));
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault) == cudaSuccess) pinned_.push_back((void*)p);
    else cudaGetLastError();
}
void AgDevice::unpin_all() {
    if (pinned_.empty()) return;
    cudaSetDevice(dev_);
    for (void* p : pinned_) cudaHostUnregister(p);
    pinned_.clear();
}
void AgDevice::timer_start() {
    CK(cudaSetDevice(dev_));
    if (!ev0_) { cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); ev0_ = a; ev1_ = b; }
    CK(cudaStreamSynchronize(m_->st));
    CK(cudaEventRecord((cudaEvent_t)ev0_, m_->st));
}
float AgDevice::timer_stop() {
    CK(cudaSetDevice(dev_));
    CK(cudaEventRecord((cudaEvent_t)ev1_, m_->st));
    CK(cudaEventSynchronize((cudaEvent_t)ev1_));
    float ms = 0; CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev0_, (cudaEvent_t)ev1_));
    return ms;
}
void AgDevice::sync() { CK(cudaSetDevice(dev_)); if (st2_) CK(cudaStreamSynchronize((cudaStream_t)st2_)); CK(cudaStreamSynchronize(m_->st)); }

void AgDevice::set_reads(const u32* bases, const u32* nmask, const uint16_t* len, u64 n_pairs, u32 stride2, u32 stridem, bool on_device) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    m.win_lo = 0; m.win_hi = ~0ull;
    m.n_pairs = n_pairs;
    if (on_device) {
        m.reads.bases = bases; m.reads.nmask = nmask; m.reads.len = len; m.reads_owned = false;
    } else {
        Timer tm(m.st);
        m.r_bases.ensure(2 * n_pairs * stride2 + 1); m.r_nmask.ensure(2 * n_pairs * stridem + 1); m.r_len.ensure(n_pairs + 1);
        CK(cudaMemcpyAsync(m.r_bases.p, bases, 2 * n_pairs * stride2 * sizeof(u32), cudaMemcpyHostToDevice, m.st));
        CK(cudaMemcpyAsync(m.r_nmask.p, nmask, 2 * n_pairs * stridem * sizeof(u32), cudaMemcpyHostToDevice, m.st));
        CK(cudaMemcpyAsync(m.r_len.p, len, n_pairs * sizeof(uint16_t), cudaMemcpyHostToDevice, m.st));
        m.reads.bases = m.r_bases.p; m.reads.nmask = m.r_nmask.p; m.reads.len = m.r_len.p; m.reads_owned = true;
        t_.h2d += tm.stop();
        t_.h2d_bytes += 2 * n_pairs * (stride2 + stridem) * sizeof(u32) + n_pairs * sizeof(uint16_t);
    }
    m.reads.stride2 = stride2; m.reads.stridem = stridem;
}

void AgDevice::set_reads_sparse(const u32* bases, const u64* exc_keys, u64 n_exc, const uint16_t* len, u64 n_pairs, u32 stride2, u32 stridem, bool overlap) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    m.n_pairs = n_pairs; m.win_lo = 0; m.win_hi = ~0ull;
    Timer tm(m.st);
    m.r_bases.ensure(2 * n_pairs * stride2 + 1); m.r_nmask.ensure(2 * n_pairs * stridem + 1); m.r_len.ensure(n_pairs + 1); m.r_exc.ensure(n_exc + 1);
    CK(cudaMemcpyAsync(m.r_len.p, len, n_pairs * sizeof(uint16_t), cudaMemcpyHostToDevice, m.st));   // k_prep needs the lengths: main stream
    cudaStream_t cs = m.st;
    if (overlap) {
        if (!st2_) { cudaStream_t s2; cudaEvent_t a, b; CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming)); st2_ = s2; ev_main_ = a; ev_reads_ = b; }
        cs = (cudaStream_t)st2_;
        CK(cudaEventRecord((cudaEvent_t)ev_main_, m.st));            // behind everything already queued on the main stream
        CK(cudaStreamWaitEvent(cs, (cudaEvent_t)ev_main_, 0));
    }
    CK(cudaMemcpyAsync(m.r_bases.p, bases, 2 * n_pairs * stride2 * sizeof(u32), cudaMemcpyHostToDevice, cs));
    CK(cudaMemsetAsync(m.r_nmask.p, 0, 2 * n_pairs * stridem * sizeof(u32), cs));
    if (n_exc) {
        CK(cudaMemcpyAsync(m.r_exc.p, exc_keys, n_exc * sizeof(u64), cudaMemcpyHostToDevice, cs));
        k_nmask_scatter<<<(unsigned)((n_exc + 255) / 256), 256, 0, cs>>>(m.r_exc.p, n_exc, m.r_nmask.p, stridem, 2 * n_pairs); launches_++;
    }
    if (overlap) { CK(cudaEventRecord((cudaEvent_t)ev_reads_, cs)); reads_pending_ = true; }
    m.reads.bases = m.r_bases.p; m.reads.nmask = m.r_nmask.p; m.reads.len = m.r_len.p; m.reads_owned = true;
    m.reads.stride2 = stride2; m.reads.stridem = stridem;
    t_.h2d += tm.stop();
    t_.h2d_bytes += 2 * n_pairs * stride2 * sizeof(u32) + n_pairs * sizeof(uint16_t) + n_exc * sizeof(u64);
}

void AgDevice::copy_reads_to_host(u32* bases, u32* nmask, uint16_t* len) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    if (st2_) CK(cudaStreamSynchronize((cudaStream_t)st2_));
    CK(cudaMemcpyAsync(bases, m.reads.bases, 2 * m.n_pairs * m.reads.stride2 * sizeof(u32), cudaMemcpyDeviceToHost, m.st));
    CK(cudaMemcpyAsync(nmask, m.reads.nmask, 2 * m.n_pairs * m.reads.stridem * sizeof(u32), cudaMemcpyDeviceToHost, m.st));
    CK(cudaMemcpyAsync(len, m.reads.len, m.n_pairs * sizeof(uint16_t), cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
}

AgDevice::ReadsView AgDevice::reads_view() {
    Impl& m = *m_;
    ReadsView v;
    v.bases = const_cast<u32*>(m.reads.bases); v.nmask = const_cast<u32*>(m.reads.nmask); v.len = const_cast<uint16_t*>(m.reads.len);
    v.n_pairs = m.n_pairs; v.stride2 = m.reads.stride2; v.stridem = m.reads.stridem;
    v.bases_bytes = 2 * m.n_pairs * m.reads.stride2 * sizeof(u32); v.nmask_bytes = 2 * m.n_pairs * m.reads.stridem * sizeof(u32); v.len_bytes = m.n_pairs * sizeof(uint16_t);
    return v;
}
AgDevice::ReadsView AgDevice::reserve_reads(u64 n_pairs, u32 stride2, u32 stridem) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    m.n_pairs = n_pairs;
    m.r_bases.ensure(2 * n_pairs * stride2 + 1); m.r_nmask.ensure(2 * n_pairs * stridem + 1); m.r_len.ensure(n_pairs + 1);
    m.reads.bases = m.r_bases.p; m.reads.nmask = m.r_nmask.p; m.reads.len = m.r_len.p; m.reads_owned = true;
    m.reads.stride2 = stride2; m.reads.stridem = stridem;
    return reads_view();
}

// ---- one broadcast of the packed reads to every GPU of the run ---------------------------------------------------------
// NCCL is bound at run time (dlopen): a single-GPU run needs no libnccl, and inside a process that already carries one (torch) the same
// copy is used.  Single process, one communicator per device (ncclCommInitAll), the three buffers as one group.
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
    NcclApi() {
        if (getenv("AG_NO_NCCL")) return;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll"); CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart"); GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
        ok = CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast;
    }
};
NcclApi& nccl() { static NcclApi a; return a; }
}  // namespace

const char* ag_device_broadcast_reads(AgDevice** devs, int n, double* seconds, size_t* bytes) {
    if (seconds) *seconds = 0;
    if (bytes) *bytes = 0;
    if (n <= 1) return "none";
    devs[0]->sync();
    const AgDevice::ReadsView src = devs[0]->reads_view();
    std::vector<AgDevice::ReadsView> dst((size_t)n);
    dst[0] = src;
    for (int i = 1; i < n; i++) dst[i] = devs[i]->reserve_reads(src.n_pairs, src.stride2, src.stridem);
    // leaders = the first context of every distinct device: one collective among them, then device-local copies to the other contexts of a device
    std::vector<int> leader((size_t)n, -1), leaders;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < i && leader[i] < 0; j++) if (devs[i]->device() == devs[j]->device()) leader[i] = leader[j] < 0 ? j : leader[j];
        if (leader[i] < 0) leaders.push_back(i);
    }
    const int nl = (int)leaders.size();
    const auto t0 = std::chrono::steady_clock::now();
    const char* how = nl > 1 ? "peer-copy" : "device-copy";
    bool done = nl <= 1;
    if (!done && nccl().ok) {
        std::vector<int> ids((size_t)nl); for (int i = 0; i < nl; i++) ids[i] = devs[leaders[i]]->device();
        std::vector<ncclComm_t> comm((size_t)nl);
        if (nccl().CommInitAll(comm.data(), nl, ids.data()) == ncclSuccess) {
            bool ok = true;
            const void* sp[3] = {src.bases, src.nmask, src.len}; const size_t nb[3] = {src.bases_bytes, src.nmask_bytes, src.len_bytes};
            for (int b = 0; b < 3 && ok; b++) {
                ok = nccl().GroupStart() == ncclSuccess;
                for (int i = 0; i < nl && ok; i++) {
                    const int c = leaders[i];
                    CK(cudaSetDevice(devs[c]->device()));
                    void* rp = b == 0 ? (void*)dst[c].bases : b == 1 ? (void*)dst[c].nmask : (void*)dst[c].len;
                    ok = nccl().Broadcast(sp[b], rp, nb[b], ncclUint8, 0, comm[i], (cudaStream_t)devs[c]->stream()) == ncclSuccess;
                }
                ok = nccl().GroupEnd() == ncclSuccess && ok;
            }
            for (int i = 0; i < nl; i++) devs[leaders[i]]->sync();
            for (int i = 0; i < nl; i++) nccl().CommDestroy(comm[i]);
            if (ok) { done = true; how = "nccl"; }
        }
    }
    if (!done) {
        for (int i = 1; i < nl; i++) {
            const int c = leaders[i];
            CK(cudaSetDevice(devs[c]->device()));
            cudaStream_t st = (cudaStream_t)devs[c]->stream();
            CK(cudaMemcpyPeerAsync(dst[c].bases, devs[c]->device(), src.bases, devs[0]->device(), src.bases_bytes, st));
            CK(cudaMemcpyPeerAsync(dst[c].nmask, devs[c]->device(), src.nmask, devs[0]->device(), src.nmask_bytes, st));
            CK(cudaMemcpyPeerAsync(dst[c].len, devs[c]->device(), src.len, devs[0]->device(), src.len_bytes, st));
        }
        for (int i = 1; i < nl; i++) devs[leaders[i]]->sync();
    }
    for (int i = 0; i < n; i++) {   // the other contexts of a device: copies inside its memory
        if (leader[i] < 0) continue;
        const int c = leader[i];
        CK(cudaSetDevice(devs[i]->device()));
        cudaStream_t st = (cudaStream_t)devs[i]->stream();
        CK(cudaMemcpyAsync(dst[i].bases, dst[c].bases, src.bases_bytes, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst[i].nmask, dst[c].nmask, src.nmask_bytes, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst[i].len, dst[c].len, src.len_bytes, cudaMemcpyDeviceToDevice, st));
    }
    for (int i = 0; i < n; i++) if (leader[i] >= 0) devs[i]->sync();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (bytes) *bytes = (src.bases_bytes + src.nmask_bytes + src.len_bytes) * (size_t)(n - 1);
    return how;
}

// ---------------------------------------------------------------------------------------------------------------------------
// GPU-side text ingestion (kernels: ag_ingest.cuh)
// ---------------------------------------------------------------------------------------------------------------------------
namespace {
struct Fd { int fd; explicit Fd(const std::string& p) : fd(open(p.c_str(), O_RDONLY)) {} ~Fd() { if (fd >= 0) close(fd); } size_t size() const { struct stat st; return fstat(fd, &st) == 0 ? (size_t)st.st_size : 0; } };
}

// newline index of text[0, len) (device, 16-byte aligned, zero-padded to a multiple of 16): positions into nl[nl_base ..]; returns the
// number of lines.  `count_only` leaves nl untouched (the caller sizes it first).  One blocking read-back of the total.
static bool ing_fallback(const char* what, int where) {   // AG_DEBUG_INGEST=1: say why a file was handed to the host parser
    static const bool dbg = getenv("AG_DEBUG_INGEST") != nullptr;
    if (dbg) fprintf(stderr, "[ag ingest] %s: not the well-formed layout (check %d) -> host parser\n", what, where);
    return false;
}

// ---- a window of the read file.  tmp/_reads.fa is written by formalizeInput (AG:3455-3471) with the pair index as the id of both
// mates, so record 2p / 2p + 1 carry ">p": the byte range of a pair-id window can be found by bisection on the file, without reading it.
// Everything is verified (ids at both ends, the record count of the window) and anything unexpected means "no window".
namespace {
// offset of the first record start ('>' at a line start) at or after `from`, its id; n = file size.  false: none / not an integer id
bool next_record(int fd, size_t n, size_t from, size_t& pos, long long& id) {
    std::vector<char> buf;
    for (size_t o = from; o < n;) {
        const size_t w = std::min<size_t>((size_t)1 << 12, n - o);
        buf.resize(w + 1);
        size_t have = 0;
        char prev = '\n';
        if (o > 0) { if (pread(fd, &prev, 1, (off_t)(o - 1)) != 1) return false; }
        if (pread(fd, buf.data(), w, (off_t)o) != (ssize_t)w) return false;
        for (size_t i = 0; i < w; i++) {
            const char before = i ? buf[i - 1] : prev;
            if (buf[i] == '>' && before == '\n') {
                char hd[32]; const size_t hw = std::min<size_t>(sizeof hd, n - (o + i));
                if (pread(fd, hd, hw, (off_t)(o + i)) != (ssize_t)hw) return false;
                long long v = 0; size_t k = 1; bool any = false;
                while (k < hw && hd[k] >= '0' && hd[k] <= '9') { v = v * 10 + (hd[k] - '0'); k++; any = true; }
                if (!any || k >= hw || hd[k] != '\n') return false;
                pos = o + i; id = v; return true;
            }
        }
        (void)have;
        o += w;
    }
    return false;
}
// offset of the first record whose id is >= target (n when there is none); ids non-decreasing along the file
bool first_record_with_id(int fd, size_t n, long long target, size_t& pos) {
    size_t lo = 0, hi = n;   // smallest offset `o` such that the first record at or after o has id >= target (or there is none)
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        size_t p; long long id;
        const bool found = next_record(fd, n, mid, p, id);
        if (!found) {   // no record start at or after mid (or a malformed header: the caller's verification catches that)
            hi = mid;
        } else if (id >= target) hi = mid; else lo = p + 1;
    }
    size_t p; long long id;
    if (lo >= n || !next_record(fd, n, lo, p, id)) { pos = n; return true; }
    pos = p; return true;
}
}  // namespace

bool AgDevice::ingest_reads(const std::string& path, AgReads& host, long long win_lo, long long win_hi) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    Fd f(path);
    if (f.fd < 0) throw AgError{"CANNOT OPEN FILE!"};
    const size_t n = f.size();
    if (n < 4) return ing_fallback("reads", 1);
    char c0 = 0, cl = 0;
    if (pread(f.fd, &c0, 1, 0) != 1 || pread(f.fd, &cl, 1, (off_t)(n - 1)) != 1 || c0 != '>' || cl != '\n') return ing_fallback("reads", 2);
    // ---- optional pair-id window [win_lo, win_hi]: only that byte range of the file is staged and packed (at its global record indices) ----
    size_t b0 = 0, b1 = n; u64 rec_base = 0, total_pairs = 0; bool windowed = false;
    if (win_lo >= 0 && win_hi >= win_lo) {
        size_t p0, pl; long long id0, idl;
        bool ok = next_record(f.fd, n, 0, p0, id0) && p0 == 0 && id0 == 0;
        // the file's last two records must be the two mates of pair total - 1
        long long id_last = -1, id_prev = -1;
        for (size_t w = (size_t)1 << 12; ok && id_prev < 0 && w <= ((size_t)1 << 18); w <<= 3) {   // scan the tail backwards for the last two record starts
            const size_t t0 = n > w ? n - w : 0;
            std::vector<char> tb(n - t0);
            if (pread(f.fd, tb.data(), tb.size(), (off_t)t0) != (ssize_t)tb.size()) { ok = false; break; }
            size_t found[2]; int nf = 0;
            for (size_t i = tb.size(); i-- > 0 && nf < 2;) if (tb[i] == '>' && (i ? tb[i - 1] == '\n' : t0 == 0)) found[nf++] = t0 + i;
            if (nf == 2) { size_t q; ok = next_record(f.fd, n, found[0], q, id_last) && q == found[0] && next_record(f.fd, n, found[1], q, id_prev) && q == found[1]; break; }
            if (t0 == 0) break;
        }
        if (ok) ok = id_last >= 0 && id_prev == id_last;
        if (ok) { total_pairs = (u64)id_last + 1; ok = (u64)win_hi < total_pairs; }
        if (ok) ok = first_record_with_id(f.fd, n, win_lo, b0) && first_record_with_id(f.fd, n, win_hi + 1, b1);
        if (ok) ok = b0 < b1 && next_record(f.fd, n, b0, pl, idl) && pl == b0 && idl == win_lo;
        if (ok && !(win_lo == 0 && (u64)win_hi + 1 == total_pairs)) { windowed = true; rec_base = 2 * (u64)win_lo; }
        else { b0 = 0; b1 = n; }
    }
    const size_t n_file = n;
    (void)n_file;
    Timer tm(st);
    // segments of at most 1 GB cut at record starts ("\n>"), each staged at a 16-byte aligned device offset (32-bit offsets inside a segment)
    const size_t SEG = (size_t)1 << 30;
    std::vector<size_t> cut(1, b0);
    while (b1 - cut.back() > SEG) {
        const size_t want = cut.back() + SEG, W = (size_t)1 << 20;
        std::vector<char> win(W);
        if (pread(f.fd, win.data(), W, (off_t)(want - W)) != (ssize_t)W) return ing_fallback("reads", 3);
        size_t k = W - 1;
        while (k > 0 && !(win[k] == '>' && win[k - 1] == '\n')) k--;
        if (k == 0) return ing_fallback("reads", 4);
        cut.push_back(want - W + k);
    }
    cut.push_back(b1);
    const size_t ns = cut.size() - 1;
    std::vector<size_t> doff(ns + 1, 0), blk0(ns + 1, 0), nl0(ns + 1, 0), rec0(ns + 1, 0);
    for (size_t s = 0; s < ns; s++) { doff[s + 1] = (doff[s] + (cut[s + 1] - cut[s]) + 15) / 16 * 16 + 16; blk0[s + 1] = blk0[s] + ((cut[s + 1] - cut[s]) + NL_B - 1) / NL_B + 2; }
    m.raw.ensure(doff[ns] + 64); m.nl_blk.ensure(blk0[ns] + 2); m.ing.ensure(8);
    for (size_t s = 0; s < ns; s++) {
        const size_t len = cut[s + 1] - cut[s];
        CK(cudaMemsetAsync(m.raw.p + doff[s] + len, 0, 32, st));
        m.stager.run(f.fd, cut[s], len, m.raw.p + doff[s], st, dev_);
    }
    const bool probe = getenv("AG_POST_TIMING") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    if (probe) { CK(cudaStreamSynchronize(st)); fprintf(stderr, "  [ingest reads] staged %.1f MB: host queued + copies done %.2f ms\n", (b1 - b0) / 1e6, tm.lap_ms()); tp0 = std::chrono::steady_clock::now(); }
    t_.h2d_bytes += b1 - b0;
    std::vector<u32> lines(ns, 0);
    for (size_t s = 0; s < ns; s++) { lines[s] = m.nl_index(st, launches_, m.raw.p + doff[s], cut[s + 1] - cut[s], blk0[s], 0, false); nl0[s + 1] = nl0[s] + lines[s]; if (lines[s] & 1) return ing_fallback("reads", 5); rec0[s + 1] = rec0[s] + lines[s] / 2; }
    const u64 R = rec0[ns];   // records staged (the window's, or the whole file's)
    if (windowed && R != 2 * (u64)(win_hi - win_lo + 1)) return ingest_reads(path, host, -1, -1);   // ids are not the pair indices after all: whole file
    if (R == 0 || (R & 1) || R >= 0xFFFFFFF0ull) return ing_fallback("reads", 6);
    const u64 R_total = windowed ? 2 * total_pairs : R;
    if (R_total >= 0xFFFFFFF0ull) return ing_fallback("reads", 6);
    m.nl.ensure(nl0[ns] + 2); m.rlen.ensure(R + 2);
    CK(cudaMemsetAsync(m.ing.p, 0, 8 * sizeof(u32), st));   // [0] bad flags, [1] max length, [2] exception count
    for (size_t s = 0; s < ns; s++) {
        m.nl_index(st, launches_, m.raw.p + doff[s], cut[s + 1] - cut[s], blk0[s], nl0[s], true);
        const u32 nr = lines[s] / 2;
        if (nr) { k_rd_len<<<(nr + 255) / 256, 256, 0, st>>>(m.raw.p + doff[s], m.nl.p + nl0[s], nr, m.rlen.p + rec0[s], m.ing.p + 1, (int*)m.ing.p); launches_++; }
    }
    volatile u32* hs = (volatile u32*)m.h_s.p;
    CK(cudaMemcpyAsync((void*)(hs + 32), m.ing.p, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hs[32]) return ing_fallback("reads", 7);   // multi-line / empty records, a read longer than 65,535: the sequential parser handles or reports it
    const u32 maxlen = hs[33];
    const u64 n_pairs = R_total / 2, w_pairs = R / 2, pair_base = rec_base / 2;
    u32 stride2 = (maxlen + 15) / 16, stridem = (maxlen + 31) / 32;
    if (!stride2) stride2 = stridem = 1;
    m.n_pairs = n_pairs; m.win_lo = pair_base; m.win_hi = pair_base + w_pairs - 1;
    m.r_bases.ensure(R_total * stride2 + 1); m.r_nmask.ensure(R_total * stridem + 1); m.r_len.ensure(n_pairs + 1);
    const u32 exc_cap = (u32)std::min<size_t>(std::max<size_t>((size_t)1 << 20, n / 64), (size_t)1 << 28);
    m.exc_key.ensure(exc_cap); m.exc_chr.ensure(exc_cap);
    k_rd_pairlen<<<(unsigned)((w_pairs + 255) / 256), 256, 0, st>>>(m.rlen.p, (u32)w_pairs, m.r_len.p + pair_base, (int*)m.ing.p); launches_++;
    for (size_t s = 0; s < ns; s++) {
        const u32 nr = lines[s] / 2;
        const u64 tasks = (u64)nr * stridem;
        if (tasks) { k_rd_pack<<<(unsigned)((tasks + 255) / 256), 256, 0, st>>>(m.raw.p + doff[s], m.nl.p + nl0[s], nr, stride2, stridem, rec_base + rec0[s], m.r_bases.p, m.r_nmask.p, m.exc_key.p, m.exc_chr.p, m.ing.p + 2, exc_cap); launches_++; }
    }
    host.len.assign(n_pairs, 0);   // (outside the window: not on the device either — see win_lo / win_hi)
    m.h_walks.ensure(w_pairs * sizeof(uint16_t) + 64);
    CK(cudaMemcpyAsync(m.h_walks.p, m.r_len.p + pair_base, w_pairs * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((void*)(hs + 32), m.ing.p, 3 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hs[32] & ING_PE_LEN) throw AgError{"INCONSISTENT PE FILES!"};
    const u32 n_exc = hs[34];
    if (n_exc > exc_cap) return ing_fallback("reads", 8);   // more masked characters than the list holds (reads of 'N' only, ...): host parser
    memcpy(host.len.data() + pair_base, m.h_walks.p, w_pairs * sizeof(uint16_t));
    host.win_lo = m.win_lo; host.win_hi = m.win_hi;
    host.exc.clear();
    if (n_exc) {
        std::vector<u64> keys(n_exc); std::vector<char> chr(n_exc);
        CK(cudaMemcpyAsync(keys.data(), m.exc_key.p, (size_t)n_exc * sizeof(u64), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(chr.data(), m.exc_chr.p, n_exc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        host.exc.resize(n_exc);
        for (u32 i = 0; i < n_exc; i++) host.exc[i] = {keys[i], chr[i]};
        std::sort(host.exc.begin(), host.exc.end());
    }
    host.exc_complete = true; host.n_pairs = n_pairs; host.stride2 = stride2; host.stridem = stridem;
    host.bases.resize(0); host.nmask.resize(0);   // device-only: AgDevice::copy_reads_to_host fills them on demand
    m.reads.bases = m.r_bases.p; m.reads.nmask = m.r_nmask.p; m.reads.len = m.r_len.p; m.reads_owned = true;
    m.reads.stride2 = stride2; m.reads.stridem = stridem;
    if (probe) fprintf(stderr, "  [ingest reads] kernels + read-backs %.2f ms\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count() * 1e3);
    t_.ingest_reads += tm.stop(); t_.reads_device++;
    t_.d2h_bytes += w_pairs * sizeof(uint16_t) + (size_t)n_exc * 9;
    if (windowed) t_.reads_windowed++;
    return true;
}

u32 AgDevice::Impl::nl_index(cudaStream_t st, u64& launches, const char* text, size_t len, size_t blk_base, size_t nl_base, bool fill) {
    Impl& m = *this;
    const size_t n16 = (len + 15) / 16;
    const unsigned nb = (unsigned)((len + NL_B - 1) / NL_B);
    if (!nb) return 0;
    u32* blk = m.nl_blk.p + blk_base;
    if (!fill) {
        k_nl_count<<<nb, NL_T, 0, st>>>((const uint4*)text, n16, blk); launches++;
        m.scanner.run(blk, blk, nb, st);   // in place, blk[nb] = total
        volatile u32* hs = (volatile u32*)m.h_s.p;
        CK(cudaMemcpyAsync((void*)(hs + 40), blk + nb, sizeof(u32), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return hs[40];
    }
    k_nl_fill<<<nb, NL_T, 0, st>>>((const uint4*)text, n16, blk, m.nl.p + nl_base); launches++;
    return 0;
}

bool AgDevice::ingest_sam(const std::string& path) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    Fd f(path);
    if (f.fd < 0) throw AgError{"CANNOT OPEN FILE!"};
    const size_t n = f.size();
    m.aln_ingested = false;
    auto none = [&]() { m.n_aln = 0; m.n_ext = 0; m.aln_ingested = true; m.aln.ensure(1); m.ext.ensure(1); t_.sam_device++; return true; };
    if (n == 0 || m.n_pairs == 0) return none();
    if (n >= 0xF0000000ull) return ing_fallback("sam", 101);
    // '@' header lines only at the top (counted on the host from the first bytes)
    size_t body = 0; u32 n_hdr = 0;
    {
        const size_t W = std::min<size_t>(n, (size_t)1 << 20);
        std::vector<char> head(W);
        if (pread(f.fd, head.data(), W, 0) != (ssize_t)W) return ing_fallback("sam", 102);
        while (body < W && head[body] == '@') { const char* l = (const char*)memchr(head.data() + body, '\n', W - body); if (!l) return ing_fallback("sam", 103); body = (size_t)(l - head.data()) + 1; n_hdr++; }
        if (body >= W && W < n) return ing_fallback("sam", 104);
        char cl = 0;
        if (pread(f.fd, &cl, 1, (off_t)(n - 1)) != 1 || cl != '\n') return ing_fallback("sam", 105);
    }
    if (body >= n) return none();
    Timer tm(st);
    m.raw.ensure(n + 64); m.nl_blk.ensure((n + NL_B - 1) / NL_B + 4); m.ing.ensure(8);
    CK(cudaMemsetAsync(m.raw.p + n, 0, 32, st));
    m.stager.run(f.fd, 0, n, m.raw.p, st, dev_);
    const bool probe = getenv("AG_POST_TIMING") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    if (probe) { CK(cudaStreamSynchronize(st)); fprintf(stderr, "  [ingest sam] staged %.1f MB: host queued + copies done %.2f ms\n", n / 1e6, tm.lap_ms()); tp0 = std::chrono::steady_clock::now(); }
    t_.h2d_bytes += n;
    const u32 lines = m.nl_index(st, launches_, m.raw.p, n, 0, 0, false);
    if (lines < n_hdr || ((lines - n_hdr) & 1)) return ing_fallback("sam", 106);   // odd number of records: BROKEN BOWTIE FILE territory, the host parser reports it
    const u32 n_rec = (lines - n_hdr) / 2;
    if (!n_rec) return none();
    m.nl.ensure((size_t)lines + 2);
    m.nl_index(st, launches_, m.raw.p, n, 0, 0, true);
    const u32* nl = m.nl.p + n_hdr;
    const u32 lost_cap = (u32)(m.n_pairs / 1000000 + 8);
    m.srec.ensure((size_t)n_rec + 1); m.s_keep.ensure((size_t)n_rec + 2); m.s_next.ensure((size_t)n_rec + 2); m.s_aoff.ensure((size_t)n_rec + 2); m.s_eoff.ensure((size_t)n_rec + 2); m.s_lost.ensure(lost_cap + 4);
    CK(cudaMemsetAsync(m.ing.p, 0, 8 * sizeof(u32), st));
    int* bad = (int*)m.ing.p;
    m.window_miss = false;
    k_sam_parse<<<(n_rec + 127) / 128, 128, 0, st>>>(m.raw.p, nl, (u32)body, n_rec, m.reads.len, m.n_pairs, m.win_lo, m.win_hi, m.srec.p, bad); launches_++;
    k_sam_sorted<<<(n_rec + 255) / 256, 256, 0, st>>>(m.srec.p, n_rec, bad); launches_++;
    k_sam_lost<<<1, 32, 0, st>>>(m.srec.p, n_rec, (long long)m.n_pairs, m.s_lost.p, lost_cap); launches_++;
    k_sam_survive<<<(n_rec + 255) / 256, 256, 0, st>>>(m.srec.p, n_rec, m.s_lost.p, m.reads.len, m.s_keep.p, m.s_next.p, bad); launches_++;
    m.scanner.run(m.s_keep.p, m.s_aoff.p, n_rec, st);
    m.scanner.run(m.s_next.p, m.s_eoff.p, n_rec, st);
    volatile u32* hs = (volatile u32*)m.h_s.p;
    CK(cudaMemcpyAsync((void*)(hs + 32), m.ing.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((void*)(hs + 33), m.s_aoff.p + n_rec, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((void*)(hs + 34), m.s_eoff.p + n_rec, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((void*)(hs + 35), m.s_lost.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const u32 flags = hs[32], n_aln = hs[33], n_ext = hs[34];
    if (flags & ING_WINDOW) m.window_miss = true;
    if ((flags & (ING_BAD_LAYOUT | ING_BAD_RECORD | ING_BAD_ORDER | ING_WINDOW)) || hs[35] > lost_cap) return ing_fallback("sam", 107);   // not the well-formed layout: host parser
    if (flags & ING_STRAND) throw AgError{"BOWTIE ALIGNMENT ERROR"};
    m.aln.ensure((size_t)n_aln + 1); m.ext.ensure((size_t)n_ext + 1);
    k_sam_fill<<<(n_rec + 127) / 128, 128, 0, st>>>(m.raw.p, nl, (u32)body, n_rec, m.reads.len, m.n_pairs, m.win_lo, m.win_hi, m.srec.p, m.s_keep.p, m.s_aoff.p, n_ext ? m.s_eoff.p : nullptr, m.aln.p, m.ext.p); launches_++;
    m.n_aln = n_aln; m.n_ext = n_ext; m.aln_ingested = true;
    if (probe) fprintf(stderr, "  [ingest sam] kernels + read-backs %.2f ms\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count() * 1e3);
    t_.ingest_sam += tm.stop(); t_.sam_device++;
    return true;
}
bool AgDevice::coverage_pileup(const std::string& path, const std::vector<u32>& chunk_len, std::vector<int>& cov) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    std::vector<u64> off(chunk_len.size() + 1, 0);
    for (size_t i = 0; i < chunk_len.size(); i++) off[i + 1] = off[i] + chunk_len[i];
    const u64 nb = off.back();
    cov.assign(nb, 0);
    Fd f(path);
    if (f.fd < 0) throw AgError{"CANNOT OPEN FILE!"};
    const size_t n = f.size();
    if (n == 0 || nb == 0) return true;
    if (n >= 0xF0000000ull || nb >= 0xFFFFFFF0ull) return ing_fallback("pileup", 201);
    size_t body = 0; u32 n_hdr = 0;
    {
        const size_t W = std::min<size_t>(n, (size_t)4 << 20);
        std::vector<char> head(W);
        if (pread(f.fd, head.data(), W, 0) != (ssize_t)W) return ing_fallback("pileup", 202);
        while (body < W && head[body] == '@') { const char* l = (const char*)memchr(head.data() + body, '\n', W - body); if (!l) return ing_fallback("pileup", 203); body = (size_t)(l - head.data()) + 1; n_hdr++; }
        if (body >= W && W < n) return ing_fallback("pileup", 204);
        char cl = 0;
        if (pread(f.fd, &cl, 1, (off_t)(n - 1)) != 1 || cl != '\n') return ing_fallback("pileup", 205);
    }
    if (body >= n) return true;
    m.raw.ensure(n + 64); m.nl_blk.ensure((n + NL_B - 1) / NL_B + 4); m.ing.ensure(8);
    CK(cudaMemsetAsync(m.raw.p + n, 0, 32, st));
    m.stager.run(f.fd, 0, n, m.raw.p, st, dev_);
    const u32 lines = m.nl_index(st, launches_, m.raw.p, n, 0, 0, false);
    if (lines < n_hdr || ((lines - n_hdr) & 1)) return ing_fallback("pileup", 206);
    const u32 n_rec = (lines - n_hdr) / 2;
    if (!n_rec) return true;
    m.nl.ensure((size_t)lines + 2);
    m.nl_index(st, launches_, m.raw.p, n, 0, 0, true);
    DBuf<u64> d_off; DBuf<u32> diff, run;
    d_off.ensure(off.size()); diff.ensure(nb + 2); run.ensure(nb + 3);
    CK(cudaMemcpyAsync(d_off.p, off.data(), off.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(diff.p, 0, (nb + 2) * sizeof(u32), st));
    CK(cudaMemsetAsync(m.ing.p, 0, 8 * sizeof(u32), st));
    k_cov_marks<<<(n_rec + 127) / 128, 128, 0, st>>>(m.raw.p, m.nl.p + n_hdr, (u32)body, n_rec, d_off.p, (u32)chunk_len.size(), diff.p, (int*)m.ing.p); launches_++;
    m.scanner.run(diff.p, run.p, nb + 1, st);   // exclusive: run[i + 1] = coverage of base i (mod 2^32, i.e. exact)
    int flags = 0;
    CK(cudaMemcpyAsync(&flags, m.ing.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cov.data(), run.p + 1, nb * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    d_off.release(); diff.release(); run.release();
    if (flags) { cov.assign(nb, 0); return ing_fallback("pileup", 207); }
    // an interval that crosses a chunk end would leak into the next chunk: intervals are clipped to their chunk in the kernel, so it cannot
    return true;
}
void AgDevice::verify_placements(const AgSeqSet& db, const AgSeqSet& qs, const std::vector<AgPlacement>& cand, std::vector<u32>& match) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    match.assign(cand.size(), 0);
    if (cand.empty()) return;
    DBuf<char> d_db, d_q; DBuf<u64> d_dbo, d_qo; DBuf<ag_place> d_c; DBuf<u32> d_m;
    d_db.ensure(db.blob.size() + 1); d_q.ensure(qs.blob.size() + 1); d_dbo.ensure(db.off.size()); d_qo.ensure(qs.off.size()); d_c.ensure(cand.size()); d_m.ensure(cand.size());
    std::vector<ag_place> hc(cand.size());
    for (size_t i = 0; i < cand.size(); i++) { hc[i].q = cand[i].q; hc[i].strand = cand[i].strand; hc[i].t = cand[i].t; hc[i].pad = 0; hc[i].start = cand[i].start; }
    CK(cudaMemcpyAsync(d_db.p, db.blob.data(), db.blob.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_q.p, qs.blob.data(), qs.blob.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_dbo.p, db.off.data(), db.off.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_qo.p, qs.off.data(), qs.off.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_c.p, hc.data(), hc.size() * sizeof(ag_place), cudaMemcpyHostToDevice, st));
    const u32 nc = (u32)cand.size();
    k_verify_placements<<<(unsigned)(((size_t)nc * 32 + 255) / 256), 256, 0, st>>>(d_db.p, d_dbo.p, d_q.p, d_qo.p, d_c.p, nc, d_m.p); launches_++;
    CK(cudaMemcpyAsync(match.data(), d_m.p, (size_t)nc * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    d_db.release(); d_q.release(); d_dbo.release(); d_qo.release(); d_c.release(); d_m.release();
}
bool AgDevice::sam_window_miss() const { return m_->window_miss; }
void AgDevice::set_reads_window(u64 lo, u64 hi) { m_->win_lo = lo; m_->win_hi = hi; }
u64 AgDevice::ingested_alignments() const { return m_->aln_ingested ? m_->n_aln : 0; }
void AgDevice::fetch_alignments(std::vector<ag_aln>& aln, std::vector<ag_seg>& ext) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    aln.assign(m.aln_ingested ? m.n_aln : 0, ag_aln{}); ext.assign(m.aln_ingested ? m.n_ext : 0, ag_seg{});
    if (!aln.empty()) CK(cudaMemcpyAsync(aln.data(), m.aln.p, aln.size() * sizeof(ag_aln), cudaMemcpyDeviceToHost, m.st));
    if (!ext.empty()) CK(cudaMemcpyAsync(ext.data(), m.ext.p, ext.size() * sizeof(ag_seg), cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
}

void AgDevice::load_unit(const AgUnitInput& in) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    Section sec(m.sections, m.st, &t_.h2d);
    if (in.aln_on_device && !m.aln_ingested) throw AgError{"internal: no ingested alignments on the device"};
    m.n_ref = in.n_ref; m.n_pos = in.n_pos; m.n_cm = in.n_cm;
    if (!in.aln_on_device) { m.n_aln = (u32)in.n_aln; m.n_ext = in.n_ext; m.aln_ingested = false; }
    if (in.n_aln >= 0xFFFFFFF0ull) throw AgError{"too many alignments for one unit"};
    const bool derive = in.threads != nullptr || in.cm_start == nullptr;   // contiMer table from the contig threads (else: explicit table)
    m.ref.ensure(in.n_pos + 1); m.cm_start.ensure((size_t)in.n_pos + 2); m.cm.ensure(in.n_cm + 1); m.chain_pos.ensure(in.n_cm + 1);
    m.chain_base.ensure(in.n_cm + 1);
    if (!in.aln_on_device) { m.aln.ensure(in.n_aln + 1); m.ext.ensure(in.n_ext + 1); }
    CK(cudaMemcpyAsync(m.ref.p, in.ref, in.n_pos, cudaMemcpyHostToDevice, m.st));
    size_t table_bytes = 0;
    if (!in.aln_on_device && in.n_aln) CK(cudaMemcpyAsync(m.aln.p, in.aln, in.n_aln * sizeof(ag_aln), cudaMemcpyHostToDevice, m.st));
    if (!in.aln_on_device && in.n_ext) CK(cudaMemcpyAsync(m.ext.p, in.ext, in.n_ext * sizeof(ag_seg), cudaMemcpyHostToDevice, m.st));
    if (in.n_cm && in.cdesc) {   // run-space contig threads: chunk bases once per file version, descriptors per unit, expansion on the device
        if (m.blob_version != in.blob_version) {
            m.contig_blob.ensure(in.blob_bytes + 1);
            CK(cudaMemcpyAsync(m.contig_blob.p, in.contig_blob, in.blob_bytes, cudaMemcpyHostToDevice, m.st));
            m.blob_version = in.blob_version; table_bytes += in.blob_bytes;
        }
        m.cdesc.ensure(in.n_desc + 1); m.cruns.ensure(in.n_runs + 1);
        CK(cudaMemcpyAsync(m.cdesc.p, in.cdesc, (size_t)in.n_desc * sizeof(ag_cdesc), cudaMemcpyHostToDevice, m.st));
        CK(cudaMemcpyAsync(m.cruns.p, in.cruns, (size_t)in.n_runs * sizeof(ag_crun), cudaMemcpyHostToDevice, m.st));
        k_chain_expand<<<(in.n_cm + 255) / 256, 256, 0, m.st>>>(m.cdesc.p, in.n_desc, m.cruns.p, m.contig_blob.p, (const unsigned char*)m.ref.p, in.n_ref, in.n_cm, m.chain_pos.p, (unsigned char*)m.chain_base.p); launches_++;
        table_bytes += (size_t)in.n_desc * sizeof(ag_cdesc) + (size_t)in.n_runs * sizeof(ag_crun);
    } else if (in.n_cm) {
        CK(cudaMemcpyAsync(m.chain_pos.p, in.chain_pos, (size_t)in.n_cm * sizeof(u32), cudaMemcpyHostToDevice, m.st));
        CK(cudaMemcpyAsync(m.chain_base.p, in.chain_base, in.n_cm, cudaMemcpyHostToDevice, m.st));
        table_bytes += (size_t)in.n_cm * 5;
    }
    if (!derive) {
        CK(cudaMemcpyAsync(m.cm_start.p, in.cm_start, ((size_t)in.n_pos + 1) * sizeof(u32), cudaMemcpyHostToDevice, m.st));
        if (in.n_cm) CK(cudaMemcpyAsync(m.cm.p, in.cm, (size_t)in.n_cm * sizeof(ag_cm), cudaMemcpyHostToDevice, m.st));
        table_bytes = ((size_t)in.n_pos + 1) * 4 + (size_t)in.n_cm * sizeof(ag_cm);
    } else {
        m.cthreads.ensure(in.n_threads + 1); m.many.ensure((size_t)in.n_pos + 2);
        if (in.n_threads) CK(cudaMemcpyAsync(m.cthreads.p, in.threads, (size_t)in.n_threads * sizeof(ag_cthread), cudaMemcpyHostToDevice, m.st));
        table_bytes = (size_t)in.n_threads * sizeof(ag_cthread);
        CK(cudaMemsetAsync(m.err_load.p, 0, sizeof(int), m.st));
        CK(cudaMemsetAsync(m.many.p, 0, ((size_t)in.n_pos + 1) * sizeof(u32), m.st));
        if (in.n_cm) { k_cm_count<<<(in.n_cm + 255) / 256, 256, 0, m.st>>>(m.chain_pos.p, in.n_cm, in.n_pos, m.many.p, m.err_load.p); launches_++; }
        m.scanner.run(m.many.p, m.cm_start.p, in.n_pos, m.st);
        if (in.n_cm) {
            CK(cudaMemsetAsync(m.many.p, 0, ((size_t)in.n_pos + 1) * sizeof(u32), m.st));
            k_cm_fill<<<(in.n_cm + 255) / 256, 256, 0, m.st>>>(m.chain_pos.p, in.n_cm, m.cthreads.p, in.n_threads, m.cm_start.p, m.many.p, m.cm.p, m.err_load.p); launches_++;
            k_cm_sort<<<(in.n_pos + 255) / 256, 256, 0, m.st>>>(m.cm_start.p, in.n_pos, m.cm.p); launches_++;
        }
        // (an inconsistent thread list sets E_CM in the upload's error word: reported at the step's synchronisation point)
    }
    t_.h2d_bytes += in.n_pos + table_bytes + (in.aln_on_device ? 0 : in.n_aln * sizeof(ag_aln) + in.n_ext * sizeof(ag_seg));
}

// ---- the step: everything from the uploaded unit to the walk records is QUEUED without a single host round trip; finish() is the one
// synchronisation point.  Buffer sizes the device alone knows (tile keys, nodes, overflow pools, start candidates, walk records) are
// capacities with an overflow bit in the step's error word: finish() then grows the capacity and the step is queued again. ----------
void AgDevice::build() { enqueue_build(); }

void AgDevice::enqueue_build() {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    cudaStream_t st = m.st;
    const u32 nA = m.n_aln, n_ref = m.n_ref, n_pos = m.n_pos;
    m.n_tiles = (n_ref + AG_TPOS - 1) / AG_TPOS;
    if (m.unit_n_ref != n_ref || m.unit_n_aln != nA) {   // a new unit: capacities from its size (kept when the same unit is built again)
        m.unit_n_ref = n_ref; m.unit_n_aln = nA;
        if (!m.key_cap_hooked) m.key_cap = std::max<u32>(m.key_cap, 2 * nA + 4096);
        m.node_cap = m.node_cap_hook ? m.node_cap_hook : std::max<u32>(m.node_cap, std::max<u32>(1u << 20, n_ref / 4 * 9 + (1u << 16)));   // 2.25 nodes per position (1.6 measured at 50x); doubled on overflow
        if (m.ovf_cap_hook) m.ovf_cap = m.ovf_cap_hook;
    }
    if (!m.node_cap) m.node_cap = std::max<u32>(1u << 20, n_ref / 4 * 9 + (1u << 16));
    if (!m.key_cap) m.key_cap = 2 * nA + 4096;
    if (!m.ovf_cap) m.ovf_cap = std::max<u32>(1u << 18, n_ref / 8);
    DevView& d = m.view;
    d = DevView{};
    d.reads = m.reads; d.ref = m.ref.p; d.n_ref = n_ref; d.n_pos = n_pos;
    d.cmt.start = m.cm_start.p; d.cmt.cm = m.cm.p; d.chain_pos = m.chain_pos.p; d.chain_base = m.chain_base.p;
    d.aln = m.aln.p; d.n_aln = nA; d.ext = m.ext.p; d.k = k_; d.iv = iv_; d.coverage = cov_;
    m.alnp.ensure(nA + 1); m.fast.ensure(nA + 1); m.ntiles.ensure(nA + 1); m.key_off.ensure((size_t)nA + 2);
    m.tile_cnt.ensure(m.n_tiles + 2); m.tile_start.ensure(m.n_tiles + 2); m.tile_flag.ensure(m.n_tiles + 2);
    m.pos_node.ensure((size_t)n_pos + 2);
    d.alnp = m.alnp.p; d.fast = m.fast.p; d.ntiles = m.ntiles.p; d.key_off = m.key_off.p;
    d.tile_cnt = m.tile_cnt.p; d.tile_start = m.tile_start.p; d.n_tiles = m.n_tiles; d.tile_flag = m.tile_flag.p;
    d.pos_node = m.pos_node.p;
    d.err = m.err.p; d.err_load = m.err_load.p;
    d.nk_ptr = m.key_off.p + nA; d.key_cap = m.key_cap; d.nn_ptr = m.counters.p + 0;
    CK(cudaMemsetAsync(m.err.p, 0, sizeof(int), st));
    CK(cudaMemsetAsync(m.counters.p, 0, 8 * sizeof(u32), st));
    CK(cudaMemsetAsync(m.tile_cnt.p, 0, (m.n_tiles + 1) * sizeof(u32), st));

    m.cm1.ensure((size_t)n_pos + 2); d.cm1 = m.cm1.p; m.pos_term.ensure((size_t)n_pos + 2); d.pos_term = m.pos_term.p;
    m.many.ensure((size_t)n_pos + 2); m.many_prefix.ensure((size_t)n_pos + 2); d.many_prefix = m.many_prefix.p;
    m.brk.ensure((size_t)n_pos + 2); m.lin_prefix.ensure((size_t)n_pos + 2); d.lin_prefix = m.lin_prefix.p;
    if (n_pos) { k_cm1<<<(n_pos + 255) / 256, 256, 0, st>>>(d, m.cm1.p, m.pos_term.p, m.many.p, m.brk.p); launches_++; }
    m.scanner.run(m.many.p, m.many_prefix.p, n_pos, st);
    m.scanner.run(m.brk.p, m.lin_prefix.p, n_pos, st);
    if (!attr_done_) { CK(cudaFuncSetAttribute(k_build, cudaFuncAttributeMaxDynamicSharedMemorySize, NODES_SMEM + NCHUNK_N * 32 * 4)); attr_done_ = true; }  // per device
#if AG_CODE4
    d.rw = (2 * m.reads.stride2 <= 32u) ? 2 * m.reads.stride2 : 0;   // words of oriented 4-bit codes per staged read (eight bases each)
#else
    d.rw = (m.reads.stride2 + m.reads.stridem <= (u32)READ_WORDS_MAX) ? m.reads.stride2 + m.reads.stridem : 0;
#endif
    // ---- prep + keys ------------------------------------------------------------------------------------------------
    m.keys.ensure((size_t)m.key_cap + 1); m.vals.ensure((size_t)m.key_cap + 1); m.keys2.ensure((size_t)m.key_cap + 1); m.vals2.ensure((size_t)m.key_cap + 1);
    d.keys = m.keys.p; d.vals = m.vals.p;
    {
        Section sec(m.sections, st, &t_.prep);
        if (nA) { k_prep<<<(nA + 255) / 256, 256, 0, st>>>(d); launches_++; }
        m.scanner.run(m.ntiles.p, m.key_off.p, nA, st);
        if (nA) { k_keys<<<(nA + 255) / 256, 256, 0, st>>>(d); launches_++; }
        m.scanner.run(m.tile_cnt.p, m.tile_start.p, m.n_tiles, st);
    }
    // ---- bucket: stable radix sort by tile (key count on the device; grids sized by the capacity) ---------------------------------------
    {
        Section sec(m.sections, st, &t_.sort);
        if (nA) {
            int bits = 1; while ((1ull << bits) < (u64)m.n_tiles) bits++;
            const unsigned nb = (m.key_cap + RS_B - 1) / RS_B;
            m.hist.ensure((size_t)RS_BINS * nb + 2);
            u32 *ka = m.keys.p, *va = m.vals.p, *kb = m.keys2.p, *vb = m.vals2.p;
            for (int shift = 0; shift < bits; shift += RS_BITS) {
                k_rs_hist<<<nb, RS_T, 0, st>>>(ka, m.hist.p, d.nk_ptr, m.key_cap, shift, nb); launches_++;
                m.scanner.run(m.hist.p, m.hist.p, (size_t)RS_BINS * nb, st, 0, 0);
                k_rs_scatter<<<nb, RS_T, 0, st>>>(ka, va, kb, vb, m.hist.p, d.nk_ptr, m.key_cap, shift, nb); launches_++;
                std::swap(ka, kb); std::swap(va, vb);
            }
            d.keys = ka; d.vals = va;
        }
    }
    // ---- nodes (+ the common-case edges as successor-item bits) -------------------------------------------------------------------
    {
        Section sec(m.sections, st, &t_.nodes);
        m.tile_base.ensure(m.n_tiles + 2); m.tile_nodes.ensure(m.n_tiles + 2); m.tile_prefix.ensure(m.n_tiles + 2); m.pos_pool.ensure((size_t)n_pos + 2);
        d.tile_base = m.tile_base.p; d.tile_nodes = m.tile_nodes.p; d.tile_prefix = m.tile_prefix.p; d.pos_pool = m.pos_pool.p;
        m.pool_c.ensure((size_t)m.node_cap + 1); m.pool_w.ensure((size_t)m.node_cap + 1); m.pool_sref.ensure(2 * (size_t)m.node_cap + 2); m.pool_pos.ensure((size_t)m.node_cap + 1);
        if (keep_counts_) m.pool_cc.ensure(6 * (size_t)m.node_cap + 6);
        m.ovf_node.ensure(m.ovf_cap); m.ovf_next.ensure(m.ovf_cap);
        d.pool_c = m.pool_c.p; d.pool_w = m.pool_w.p; d.pool_sref = m.pool_sref.p; d.pool_pos = m.pool_pos.p; d.pool_cc = keep_counts_ ? m.pool_cc.p : nullptr;
        d.node_cap = m.node_cap; d.pool_count = m.counters.p + 0;
        d.ovf.node = m.ovf_node.p; d.ovf.next = m.ovf_next.p; d.ovf.count = m.counters.p + 1; d.ovf.cap = m.ovf_cap; d.ovf.err = m.err.p;
        CK(cudaMemsetAsync(m.tile_flag.p, 0, ((size_t)m.n_tiles + 1) * sizeof(u32), st));
        if (reads_pending_) { CK(cudaStreamWaitEvent(st, (cudaEvent_t)ev_reads_, 0)); reads_pending_ = false; }   // overlapped reads upload (set_reads_sparse)
        const u32 maxlen = m.reads.stride2 * 16;
        const u32 rsw = (ST_HDR_W + (maxlen + 7) / 8 + 3) / 4 * 4;                       // staged record words: header + oriented 4-bit codes, 16-byte multiple
        const size_t smem_tma = NODES_SMEM + 2 * (size_t)ST_CH * rsw * 4;
        if (!tma_off_ && smem_tma <= 36 * 1024) {                                       // (reads of up to ~500 bases; longer ones take the per-thread staging kernel)
            if (!attr_tma_done_) { CK(cudaFuncSetAttribute(k_build_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 1024)); attr_tma_done_ = true; }
            m.stage.ensure((size_t)m.key_cap * rsw + 64);
            {
                Section sec2(m.sections, st, &t_.stage);
                k_stage<<<GS_BLOCKS * 2, GS_T, 0, st>>>(d, m.stage.p, rsw); launches_++;
            }
            Section sec3(m.sections, st, &t_.build_kernel);
            if (m.n_tiles) { k_build_tma<<<m.n_tiles, AG_TILE, smem_tma, st>>>(d, m.stage.p, rsw); launches_++; }
        } else {
            Section sec3(m.sections, st, &t_.build_kernel);
            if (m.n_tiles) { k_build<<<m.n_tiles, AG_TILE, NODES_SMEM + NCHUNK_N * d.rw * 4, st>>>(d); launches_++; }
        }
    }
    // ---- tile blocks -> position order, successor bits -> successor indices (arrays sized by the node CAPACITY) ----------------------------
    {
        Section sec(m.sections, st, &t_.finalize);
        const size_t nc = m.node_cap;
        m.node_c.ensure(nc + 1); m.node_w.ensure(nc + 1); m.node_sref.ensure(2 * nc + 2); m.node_pos.ensure(nc + 1);
        if (keep_counts_) m.node_cc.ensure(6 * nc + 6);
        d.node_c = m.node_c.p; d.node_w = m.node_w.p; d.node_sref = m.node_sref.p; d.node_pos = m.node_pos.p; d.node_cc = keep_counts_ ? m.node_cc.p : nullptr;
        m.eovf_head.ensure(nc + 1);
        if (!m.eovf_cap) m.eovf_cap = m.eovf_cap_init ? m.eovf_cap_init : std::max<u32>(1u << 18, m.node_cap / 16);
        m.eovf_target.ensure(m.eovf_cap); m.eovf_next.ensure(m.eovf_cap);
        d.eovf_head = m.eovf_head.p; d.eovf_target = m.eovf_target.p; d.eovf_next = m.eovf_next.p; d.eovf_count = m.counters.p + 2; d.eovf_cap = m.eovf_cap;
        m.scanner.run(m.tile_nodes.p, m.tile_prefix.p, m.n_tiles, st);
        k_posfix<<<(n_pos + 1 + 255) / 256, 256, 0, st>>>(d); launches_++;
        k_succ<<<GS_BLOCKS, GS_T, 0, st>>>(d); launches_++;
    }
    // ---- generic edge sweep over the flagged tiles ----------------------------------------------------------------------------------
    {
        Section sec(m.sections, st, &t_.edges);
        if (m.n_tiles) { k_edges<<<m.n_tiles, AG_TILE, 0, st>>>(d); launches_++; }
    }
    m.build_queued = true; m.walk_queued = false;
    t_.n_tiles = m.n_tiles;
}

// forced-link chains, start candidates, components, replay, compaction of the walk records into host memory — queued behind the build
void AgDevice::enqueue_walk(bool records_to_host) {
    Impl& m = *m_; cudaStream_t st = m.st; DevView& d = m.view;
    const size_t nc = m.node_cap;
    if (!m.cand_cap) m.cand_cap = (u32)std::max<size_t>(1u << 16, nc / 16);
    if (!m.hwalk_cap) m.hwalk_cap = (u32)std::max<size_t>(1u << 16, m.n_ref / 16);
    m.walk_next.ensure(nc + 1); m.parent.ensure(nc + 1); m.cmin.ensure(nc + 2); m.cmax.ensure(nc + 1);
    d.walk_next = m.walk_next.p; d.parent = m.parent.p; d.cmin = m.cmin.p; d.cmax = m.cmax.p;
    d.walk_count = m.counters.p + 3;
    {   // forced-link chains
        Section sec(m.sections, st, &t_.chains);
        m.indeg.ensure(nc + 1); m.fnext.ensure(nc + 1); m.fprev.ensure(nc + 1); d.fprev = m.fprev.p; m.chain_a.ensure(nc + 1); m.chain_b.ensure(nc + 1);
        d.indeg = m.indeg.p; d.fnext = m.fnext.p; d.chain_a = m.chain_a.p; d.chain_b = m.chain_b.p;
        k_uf_init<<<GS_BLOCKS, GS_T, 0, st>>>(d); launches_++;
        k_indeg<<<GS_BLOCKS, GS_T, 0, st>>>(d); launches_++;
        ag_chain *a = m.chain_a.p, *b = m.chain_b.p;
        k_links_rank_local<<<(unsigned)((nc + 1023) / 1024), 1024, 0, st>>>(d, a); launches_++;
        for (int round = 0; round < m.rank_rounds; round++) {  // links that leave a 1024-node block: a fixed number of global rounds, more on request (E_RANK_MORE)
            const bool last = round + 1 == m.rank_rounds;
            k_rank<<<GS_BLOCKS, GS_T, 0, st>>>(a, b, d.nn_ptr, m.err.p, last ? 1 : 0, d.node_w, last && m.scanner.one_pass ? m.indeg.p : nullptr); launches_++;
            std::swap(a, b);
        }
        d.chain = a;
        // start candidates = chain heads, compacted in node order
        m.cand_rank.ensure(nc + 2); m.cand_node.ensure((size_t)m.cand_cap + 1); m.cand_label.ensure((size_t)m.cand_cap + 1);
        d.cand_rank = m.cand_rank.p; d.cand_node = m.cand_node.p; d.cand_label = m.cand_label.p;
        d.ncand_ptr = m.cand_rank.p + nc; d.cand_cap = m.cand_cap;
        if (m.scanner.one_pass) m.scanner.run(m.indeg.p, m.cand_rank.p, nc, st, 0, 1, d.nn_ptr);   // flags written by the last ranking round; scan over the device-side node count
        else { k_cand_flag<<<GS_BLOCKS, GS_T, 0, st>>>(d, m.indeg.p); launches_++; m.scanner.run(m.indeg.p, m.cand_rank.p, nc, st); }
        k_cand_scatter<<<GS_BLOCKS, GS_T, 0, st>>>(d, m.indeg.p); launches_++;
    }
    // every walk starts at a chain head, so the candidate count bounds the number of walk records
    m.walk_cap = m.cand_cap + 1; m.walks.ensure(m.walk_cap); m.walk_used.ensure((size_t)m.cand_cap + 2);
    m.h_wrec.ensure((size_t)m.hwalk_cap * sizeof(ag_walk));
    m.walks2.ensure(m.walk_cap);
    d.walks = m.walks.p; d.walks_sorted = m.walks2.p; d.walk_cap = m.walk_cap; d.walk_used = m.walk_used.p;
    d.walks_host = (ag_walk*)m.h_wrec.p; d.hwalk_cap = m.hwalk_cap;
    CK(cudaMemsetAsync(m.walk_used.p, 0, ((size_t)m.cand_cap + 1) * sizeof(u32), st));
    {
        Section sec(m.sections, st, &t_.components);
        m.hrec.ensure(nc + 1); d.hrec = m.hrec.p; m.hdet.ensure(nc + 1); d.hdet = m.hdet.p;
        k_hrec<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d); launches_++;
        k_uf_tails<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d); launches_++;
        k_uf_flatten<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d); launches_++;
    }
    {
        Section sec(m.sections, st, &t_.walk);
        k_walk_components<<<148 * 16, 256, 0, st>>>(d, m.counters.p + 6); launches_++;   // counters[6]: next candidate (zeroed with the other counters when the build was queued)
    }
    {
        Section sec(m.sections, st, &t_.d2h);
        m.walk_rank.ensure((size_t)m.cand_cap + 2);
        m.scanner.run(m.walk_used.p, m.walk_rank.p, m.cand_cap, st);
        k_walk_compact<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d, m.walk_rank.p); launches_++;
        if (records_to_host) { k_walk_to_host<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d, m.walk_rank.p); launches_++; }   // (the fused extension only ships the emitted ones)
    }
    chains_valid_ = true;
    m.walk_queued = true;
}

// The step's synchronisation point.  Returns true when a capacity was exceeded: it has been grown and the caller queues the step again.
bool AgDevice::finish() {
    Impl& m = *m_; cudaStream_t st = m.st; DevView& d = m.view;
    if (!m.build_queued && !m.walk_queued) return false;
    k_status<<<1, 32, 0, st>>>(d, m.counters.p, m.walk_queued ? m.walk_rank.p : nullptr, m.select_queued ? m.sel_info.p : nullptr, m.status.p); launches_++;
    volatile u32* hs = (volatile u32*)m.h_s.p;
    CK(cudaMemcpyAsync((void*)hs, m.status.p, 12 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    m.sections.collect();
    const int err = (int)hs[0];
    const u32 nn = hs[1], nk = hs[4], ncand = hs[5], nw = hs[6];
    if (err & E_BAD_ALN) { m.build_queued = false; throw AgError{"BOWTIE ALIGNMENT ERROR: alignment outside the unit"}; }
    if (err & E_CM) { m.build_queued = false; throw AgError{"CONTIG ALIGNMENT ERROR: inconsistent contig threads"}; }
    bool redo = false;
    if (err & E_KEY_CAP) { m.key_cap = nk + nk / 8 + 4096; redo = true; }
    if (err & E_NODE_CAP) { if (m.node_cap >= (1u << 31)) throw AgError{"node table exhausted"}; m.node_cap = std::max<u32>(m.node_cap * 2, nn + nn / 8); redo = true; }
    if (err & E_OVF) { if (m.ovf_cap >= (1u << 30)) throw AgError{"node table exhausted"}; m.ovf_cap *= 4; redo = true; }
    if (err & E_EDGE_OVF) { if (m.eovf_cap >= (1u << 30)) throw AgError{"edge overflow pool exhausted"}; m.eovf_cap *= 4; redo = true; }
    if (err & E_CAND_CAP) { m.cand_cap = ncand + ncand / 8 + 1024; redo = true; }
    if (err & E_HWALK_CAP) { m.hwalk_cap = nw + nw / 8 + 1024; redo = true; }
    if (err & E_BASES_CAP) { m.bases_cap = hs[8] + hs[8] / 8 + (1u << 20); redo = true; }
    if (err & E_RANK_MORE) { if (m.rank_rounds >= 24) throw AgError{"internal: chain ranking did not converge"}; m.rank_rounds += 2; redo = true; }
    if (err & (E_WALK_CAP | E_MAT_CAP)) throw AgError{"walk record buffer exhausted"};
    if (redo) { t_.regrows++; m.build_queued = m.walk_queued = m.select_queued = false; occ_pending_ = false; return true; }
    m.n_nodes = nn; m.n_keys = nk; t_.n_nodes = nn; t_.n_keys = nk; t_.n_edges_ovf = hs[3];
    if (m.walk_queued) { m.n_cand = ncand; m.n_walks = nw; t_.n_components = ncand; }
    if (m.select_queued) { m.n_sel = hs[7]; m.sel_bases = hs[8]; m.sel_trigger = hs[9] != 0; }
    m.build_queued = m.walk_queued = m.select_queued = false;
    return false;
}
void AgDevice::build_sync() { while (finish()) enqueue_build(); }

// emission filter, output offsets, materialisation and the copies to page-locked host memory — queued behind the walk, no host round trip
void AgDevice::enqueue_select() {
    Impl& m = *m_; cudaStream_t st = m.st; DevView& d = m.view;
    const size_t wc = m.hwalk_cap;
    if (!m.bases_cap) m.bases_cap = (u32)std::min<u64>(0xFFFFFF00ull, (u64)m.n_ref + m.n_ref / 4 + (1u << 20));
    d.bases_cap = m.bases_cap;
    m.sel_E.ensure(wc + 2); m.sel_M.ensure(wc + 2); m.sel_flag.ensure(wc + 2); m.sel_len.ensure(wc + 2); m.sel_rank.ensure(wc + 2); m.sel_soff.ensure(wc + 2); m.sel_info.ensure(8);
    m.sel_start.ensure(wc + 1); m.sel_off.ensure(wc + 1); m.sel_tails.ensure(3 * wc + 1); m.sel_walks.ensure(wc + 1); m.sel_off32.ensure(wc + 2);
    m.out_bases.ensure((size_t)m.bases_cap + 16);
    m.h_selw.ensure(wc * sizeof(ag_walk) + 64); m.h_selo.ensure((wc + 2) * sizeof(u32) + 64); m.h_bases.ensure((size_t)m.bases_cap + 16);
    {
        Section sec(m.sections, st, &t_.select);
        k_sel_prepare<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d, m.walk_rank.p, m.sel_E.p, m.sel_info.p); launches_++;   // (also clears sel_info)
        if (m.scanner.one_pass) m.scanner.run_max(m.sel_E.p, m.sel_M.p, m.hwalk_cap, m.walk_rank.p + m.cand_cap, st);
        else { k_excl_max_scan<<<1, 1024, 0, st>>>(m.sel_E.p, m.sel_M.p, m.walk_rank.p + m.cand_cap, m.hwalk_cap); launches_++; }
        k_sel_flag<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d, m.walk_rank.p, m.sel_E.p, m.sel_M.p, m.sel_flag.p, m.sel_len.p, m.sel_info.p + 2); launches_++;
        m.scanner.run(m.sel_flag.p, m.sel_rank.p, wc, st);    // sel_rank[wc] = emitted walks
        m.scanner.run(m.sel_len.p, m.sel_soff.p, wc, st);     // sel_soff[wc] = their bases
        k_sel_fill<<<GS_BLOCKS / 4, GS_T, 0, st>>>(d, m.walk_rank.p, m.sel_flag.p, m.sel_rank.p, m.sel_soff.p, m.sel_start.p, m.sel_off.p, m.sel_tails.p, m.sel_walks.p, m.sel_off32.p, m.sel_info.p); launches_++;   // (sel_info[0..1] = emitted walks, their bases)
    }
    {   // emitted walk records + their offsets to page-locked host memory, then the early status block and its event: the host starts on the
        // contig records while the materialisation below is still running
        Section sec(m.sections, st, &t_.select);
        const u32* n_sel = m.sel_info.p + 0;
        k_to_host16<<<GS_BLOCKS / 4, GS_T, 0, st>>>((const uint4*)m.sel_walks.p, (uint4*)m.h_selw.p, n_sel, (u32)sizeof(ag_walk), 0, m.err.p); launches_++;
        k_to_host16<<<GS_BLOCKS / 4, GS_T, 0, st>>>((const uint4*)m.sel_off32.p, (uint4*)m.h_selo.p, n_sel, 4u, 1, m.err.p); launches_++;
        k_status_early<<<1, 32, 0, st>>>(d, m.sel_info.p, (volatile u32*)m.h_s.p + 16); launches_++;
        if (!m.ev_early) CK(cudaEventCreateWithFlags(&m.ev_early, cudaEventDisableTiming));
        CK(cudaEventRecord(m.ev_early, st));
    }
    {
        Section sec(m.sections, st, &t_.materialize);
        const u32* n_sel = m.sel_info.p + 0;
        const u32 cap = m.cand_cap + 1;
        m.mat_detours.ensure(cap); m.tail_end.ensure((size_t)m.node_cap + 1);
        CK(cudaMemsetAsync(m.counters.p + 4, 0, 2 * sizeof(u32), st));
        CK(cudaMemsetAsync(m.tail_end.p, 0xFF, (size_t)m.node_cap * sizeof(u32), st));
        k_mat_items<<<GS_BLOCKS / 8, 128, 0, st>>>(d, m.sel_start.p, m.sel_off.p, 0, n_sel, m.tail_end.p, m.mat_detours.p, m.counters.p + 4, cap); launches_++;
        k_mat_nodes<<<GS_BLOCKS, GS_T, 0, st>>>(d, 0, d.nn_ptr, m.tail_end.p, m.out_bases.p); launches_++;
        k_mat_detours<<<148u * 8u, 256, 0, st>>>(d, m.mat_detours.p, m.counters.p + 4, m.out_bases.p); launches_++;
        k_mat_tails<<<GS_BLOCKS / 8, 128, 0, st>>>(d, m.sel_tails.p, m.sel_off.p, 0, n_sel, m.out_bases.p); launches_++;
        // the bases to page-locked host memory (count known to the device only)
        k_to_host16<<<GS_BLOCKS, GS_T, 0, st>>>((const uint4*)m.out_bases.p, (uint4*)m.h_bases.p, m.sel_info.p + 1, 1u, 0, m.err.p); launches_++;
    }
    m.select_queued = true;
}

// extendContigs1 up to the emitted contigs as ONE queued step: walk, emission filter, materialisation, copies — and one synchronisation.
// `emitted` = the walk records that pass the emission filter, in scan order; contig i = bases[offs[i], offs[i + 1]).
bool AgDevice::extend_emitted(std::vector<ag_walk>& emitted, char*& bases, std::vector<u64>& offs, u64& n_walks, const std::function<void()>& on_records) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    emitted.clear(); offs.assign(1, 0); bases = nullptr; n_walks = 0;
    if (!m.build_queued && !m.n_nodes) { occupancy_begin(); return false; }
    bool early = false;
    for (;;) {
        enqueue_walk(false);
        enqueue_select();
        occupancy_begin();
        early = false;
        if (on_records) {   // records + offsets of the emitted walks are in host memory before the bases: let the caller work on them meanwhile
            CK(cudaEventSynchronize(m.ev_early));
            volatile u32* he = (volatile u32*)m.h_s.p + 16;
            if (!(he[0] & E_FATAL) && !he[3]) {
                const u32 ns = he[1];
                emitted.resize(ns); offs.assign((size_t)ns + 1, 0);
                if (ns) {
                    memcpy(emitted.data(), m.h_selw.p, (size_t)ns * sizeof(ag_walk));
                    const u32* o32 = (const u32*)m.h_selo.p;
                    for (u32 i = 0; i <= ns; i++) offs[i] = o32[i];
                }
                bases = (char*)m.h_bases.p;
                on_records();
                early = true;
            }
        }
        if (!finish()) break;
        enqueue_build();
    }
    n_walks = m.n_walks; t_.n_walks = m.n_walks;
    if (m.sel_trigger) {   // a > 100 kbp contig: the reference's scan skips positions from here on (AG:2194-2202) — exact sequential replay, host emission filter
        std::vector<ag_walk> walks;
        walk_sequential();
        cudaStream_t st = m.st;
        volatile u32* hs = (volatile u32*)m.h_s.p;
        CK(cudaMemcpyAsync((void*)(hs + 0), m.counters.p + 3, sizeof(u32), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync((void*)(hs + 1), m.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const u32 nw = hs[0];
        if (hs[1]) throw AgError{"walk record buffer exhausted"};
        walks.resize(nw);
        if (nw) {
            m.h_walks.ensure((size_t)nw * sizeof(ag_walk));
            CK(cudaMemcpyAsync(m.h_walks.p, m.walks.p, (size_t)nw * sizeof(ag_walk), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            memcpy(walks.data(), m.h_walks.p, (size_t)nw * sizeof(ag_walk));
        }
        n_walks = nw; t_.n_walks = nw;
        std::vector<u32> sel;
        ag_select_emitted(walks, sel);
        materialize_begin(walks, sel, bases, offs);
        materialize_wait();
        emitted.resize(sel.size());
        for (size_t i = 0; i < sel.size(); i++) emitted[i] = walks[sel[i]];
        t_.d2h_bytes += (size_t)nw * sizeof(ag_walk);
        return false;
    }
    const u32 ns = m.n_sel;
    t_.d2h_bytes += (size_t)ns * sizeof(ag_walk) + m.sel_bases;
    if (early && emitted.size() == ns) return true;   // what on_records saw is what the finished step produced
    emitted.resize(ns); offs.assign((size_t)ns + 1, 0);
    if (ns) {
        memcpy(emitted.data(), m.h_selw.p, (size_t)ns * sizeof(ag_walk));
        const u32* o32 = (const u32*)m.h_selo.p;
        for (u32 i = 0; i <= ns; i++) offs[i] = o32[i];
    }
    bases = (char*)m.h_bases.p;
    return false;
}

void AgDevice::walk_sequential() {
    Impl& m = *m_; cudaStream_t st = m.st; DevView& d = m.view; u32 nn = m.n_nodes;
    Timer tm(st);
    // reset marks to the coverage filter state and replay in one thread
    m.msuf.ensure(nn + 1); m.mnode.ensure(nn + 1); d.msuf = m.msuf.p; d.mnode = m.mnode.p;
    // with the skip rule a walk may start inside a chain: any live node can be a start
    m.walk_cap = nn + 1; m.walks.ensure(m.walk_cap);
    d.walks = m.walks.p; d.walk_cap = m.walk_cap;
    k_reset_marks<<<(nn + 255) / 256, 256, 0, st>>>(d, nn); launches_++;
    chains_valid_ = false;  // chains stay valid as data; materialisation switches to the STOP-bit rule
    CK(cudaMemsetAsync(m.counters.p + 3, 0, sizeof(u32), st));
    k_walk_sequential<<<1, 32, 0, st>>>(d); launches_++;
    t_.walk += tm.stop();
    t_.walk_fallback = 1;
}

void AgDevice::extend(std::vector<ag_walk>& walks) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    walks.clear();
    if (!m.build_queued && !m.n_nodes) return;   // (checked build of an empty unit)
    for (;;) {
        enqueue_walk();                 // behind the build, which may still be in flight
        if (!finish()) break;           // ONE blocking synchronisation for the whole build + walk
        enqueue_build();                // a capacity was exceeded: everything again with the larger one
    }
    if (!m.n_nodes) return;
    walks.resize(m.n_walks);
    if (m.n_walks) memcpy(walks.data(), m.h_wrec.p, (size_t)m.n_walks * sizeof(ag_walk));   // written by k_walk_compact straight into page-locked host memory
    t_.d2h_bytes += (size_t)m.n_walks * sizeof(ag_walk);
    // the 1000-position skip of the reference's scan (AG:2194-2202) only matters once a contig longer than 100 kbp has been
    // emitted; detect that on the emitted sequence and, if so, replay sequentially on the device (exact, slow, rare)
    {
        u32 bso = AG_NONE, beo = AG_NONE; bool have = false, trigger = false;
        for (size_t i = 0; i < walks.size() && !trigger; i++) {
            const ag_walk& r = walks[i];
            u32 eoff = r.eoff;
            if (((r.flags >> 1) & 3) == 0) eoff = eoff + (r.tail_soff_len >> 16) - 1;
            bool contained = have && bso <= r.soff && beo >= eoff;
            if (!contained) { bso = r.soff; beo = eoff; have = true; if (beo - bso > 100000u) trigger = true; }
        }
        if (trigger) {
            walk_sequential();
            volatile u32* hs = (volatile u32*)m.h_s.p;
            CK(cudaMemcpyAsync((void*)(hs + 0), m.counters.p + 3, sizeof(u32), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync((void*)(hs + 1), m.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const u32 nw = hs[0];
            if (hs[1]) throw AgError{"walk record buffer exhausted"};
            walks.resize(nw);
            if (nw) {
                m.h_walks.ensure((size_t)nw * sizeof(ag_walk));
                CK(cudaMemcpyAsync(m.h_walks.p, m.walks.p, (size_t)nw * sizeof(ag_walk), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                memcpy(walks.data(), m.h_walks.p, (size_t)nw * sizeof(ag_walk));
            }
            t_.d2h_bytes += (size_t)nw * sizeof(ag_walk);
        }
    }
    t_.n_walks = walks.size();
}

void AgDevice::materialize_begin(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char*& bases, std::vector<u64>& offs) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st; DevView& d = m.view;
    offs.assign(sel.size() + 1, 0);
    bases = nullptr;
    if (sel.empty()) return;
    m.h_sel.ensure(sel.size() * (sizeof(u64) + 4 * sizeof(u32)) + 64);   // page-locked staging: offsets, start nodes, tail descriptors
    u64* h_off = (u64*)m.h_sel.p; u32* h_start = (u32*)(h_off + sel.size()); u32* h_tails = h_start + sel.size();
    for (size_t i = 0; i < sel.size(); i++) {
        const ag_walk& r = walks[sel[i]];
        u32 tl = ag_walk_tail_len(r);
        h_start[i] = r.start_node; offs[i + 1] = offs[i] + r.len + tl; h_off[i] = offs[i];
        h_tails[3 * i] = r.tail_sread; h_tails[3 * i + 1] = tl ? r.tail_soff_len : 0; h_tails[3 * i + 2] = r.len;
    }
    if (offs.back() >= 0xFFFFFFF0ull) throw AgError{"materialise: more than 4 GB of contig bases in one unit"};
    if (!ev_mat0_) { cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); ev_mat0_ = a; ev_mat1_ = b; }
    CK(cudaEventRecord((cudaEvent_t)ev_mat0_, st));
    m.sel_start.ensure(sel.size() + 1); m.sel_off.ensure(sel.size() + 1); m.out_bases.ensure(offs.back() + 1); m.sel_tails.ensure(3 * sel.size() + 1);
    CK(cudaMemcpyAsync(m.sel_off.p, h_off, sel.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m.sel_start.p, h_start, sel.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m.sel_tails.p, h_tails, 3 * sel.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
    if (chains_valid_) {
        const u32 cap = m.n_cand + 1, nn = m.n_nodes;
        m.mat_detours.ensure(cap); m.tail_end.ensure((size_t)nn + 1);
        CK(cudaMemsetAsync(m.counters.p + 4, 0, 2 * sizeof(u32), st));
        CK(cudaMemsetAsync(m.tail_end.p, 0xFF, (size_t)nn * sizeof(u32), st));
        k_mat_items<<<((u32)sel.size() + 127) / 128, 128, 0, st>>>(d, m.sel_start.p, m.sel_off.p, (u32)sel.size(), nullptr, m.tail_end.p, m.mat_detours.p, m.counters.p + 4, cap); launches_++;
        k_mat_nodes<<<(nn + 255) / 256, 256, 0, st>>>(d, nn, nullptr, m.tail_end.p, m.out_bases.p); launches_++;
        k_mat_detours<<<std::min<u32>(cap, 148u * 8u), 256, 0, st>>>(d, m.mat_detours.p, m.counters.p + 4, m.out_bases.p); launches_++;
    } else {
        k_materialize_seq<<<((u32)sel.size() + 127) / 128, 128, 0, st>>>(d, m.sel_start.p, m.sel_off.p, (u32)sel.size(), m.out_bases.p); launches_++;
    }
    k_mat_tails<<<((u32)sel.size() + 127) / 128, 128, 0, st>>>(d, m.sel_tails.p, m.sel_off.p, (u32)sel.size(), nullptr, m.out_bases.p); launches_++;
    m.h_bases.ensure(offs.back());
    CK(cudaMemcpyAsync(m.h_bases.p, m.out_bases.p, offs.back(), cudaMemcpyDeviceToHost, st));
    volatile u32* hs = (volatile u32*)m.h_s.p;
    CK(cudaMemcpyAsync((void*)(hs + 16), m.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord((cudaEvent_t)ev_mat1_, st));
    bases = (char*)m.h_bases.p;   // the post passes read (and patch) the page-locked buffer in place — after materialize_wait()
    mat_pending_ = true; mat_bytes_ = offs.back();
    t_.h2d_bytes += sel.size() * 24; t_.d2h_bytes += offs.back();
}

void AgDevice::materialize_wait() {
    if (!mat_pending_) return;
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    mat_pending_ = false;
    CK(cudaEventSynchronize((cudaEvent_t)ev_mat1_));
    float ms = 0; CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev_mat0_, (cudaEvent_t)ev_mat1_));
    t_.materialize += ms;
    if (((volatile u32*)m.h_s.p)[16]) throw AgError{"materialise item buffer exhausted"};
}

void AgDevice::occupancy_begin() {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_; cudaStream_t st = m.st;
    const size_t nb = ((size_t)m.n_pos + 7) / 8;
    occ_pending_ = true;
    if (!nb) return;
    m.occ.ensure(nb + 1);
    k_occupancy<<<((u32)nb + 255) / 256, 256, 0, st>>>(m.view, m.occ.p); launches_++;
    m.h_occ.ensure(nb);
    CK(cudaMemcpyAsync(m.h_occ.p, m.occ.p, nb, cudaMemcpyDeviceToHost, st));
    t_.d2h_bytes += nb;
}
void AgDevice::occupancy_wait(std::vector<unsigned char>& bits) {
    CK(cudaSetDevice(dev_));
    Impl& m = *m_;
    if (!occ_pending_) occupancy_begin();
    occ_pending_ = false;
    const size_t nb = ((size_t)m.n_pos + 7) / 8;
    bits.assign(nb, 0);
    if (!nb) return;
    CK(cudaStreamSynchronize(m.st));
    memcpy(bits.data(), m.h_occ.p, nb);
}

void AgDevice::dump_nodes(AgNodeDump& dd) {
    CK(cudaSetDevice(dev_));
    build_sync();
    Impl& m = *m_; cudaStream_t st = m.st; u32 nn = m.n_nodes, n_pos = m.n_pos;
    if (!keep_counts_ || !m.view.node_cc) throw AgError{"node dump needs the per-node counters: set option keep_counts before building"};
    std::vector<u32> pos_node((size_t)n_pos + 1), cc(6 * (size_t)nn), sref(2 * (size_t)nn), npos(nn);
    std::vector<ag_nodec> nc(nn);
    std::vector<ag_nodew> nw(nn);
    u32 cnt[4];
    CK(cudaMemcpyAsync(cnt, m.counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pos_node.data(), m.pos_node.p, ((size_t)n_pos + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (nn) {
        CK(cudaMemcpyAsync(cc.data(), m.node_cc.p, (size_t)nn * 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sref.data(), m.node_sref.p, (size_t)nn * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(npos.data(), m.node_pos.p, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(nc.data(), m.node_c.p, (size_t)nn * sizeof(ag_nodec), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(nw.data(), m.node_w.p, (size_t)nn * sizeof(ag_nodew), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    std::vector<u32> eh(nn), et(cnt[2]), en(cnt[2]);
    if (nn) CK(cudaMemcpyAsync(eh.data(), m.eovf_head.p, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
    if (cnt[2]) { CK(cudaMemcpyAsync(et.data(), m.eovf_target.p, (size_t)cnt[2] * 4, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(en.data(), m.eovf_next.p, (size_t)cnt[2] * 4, cudaMemcpyDeviceToHost, st)); }
    CK(cudaStreamSynchronize(st));
    dd = AgNodeDump();
    dd.edge_start.push_back(0);
    for (u32 v = 0; v < nn; v++) {
        const u32 q = npos[v];
        dd.pos.push_back(q); dd.item.push_back(v - pos_node[q]); dd.cov.push_back(cc[6 * (size_t)v]);
        for (int j = 0; j < 5; j++) dd.cnt.push_back(cc[6 * (size_t)v + 1 + j]);
        dd.cid.push_back(nc[v].cid); dd.coff.push_back(nc[v].coff); dd.cid0.push_back(nc[v].cid0); dd.coff0.push_back(nc[v].coff0); dd.moff.push_back(nw[v].moff);
        dd.sread.push_back(sref[2 * (size_t)v]); dd.soff_len.push_back(sref[2 * (size_t)v + 1]);
        std::vector<u32> e;
        if (nw[v].succ0 != AG_NONE) e.push_back(nw[v].succ0);
        if (nw[v].succ1 != AG_NONE) e.push_back(nw[v].succ1);
        if (nw[v].misc & AG_NW_OVF) for (u32 o = eh[v]; o != AG_NONE; o = en[o]) e.push_back(et[o]);
        std::sort(e.begin(), e.end());
        for (u32 x : e) dd.edge_target.push_back(x);
        dd.edge_start.push_back((u32)dd.edge_target.size());
    }
}

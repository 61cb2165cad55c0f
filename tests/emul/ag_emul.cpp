// ag_emul — HOST EMULATION of the device pipeline, for the CPU test-suite only.
//
// *** TEST INFRASTRUCTURE.  Never built into, linked with or loaded by the product. ***
// There is no GPU in the development container, so the per-thread logic of the CUDA kernels lives in
// aligngraph_b200/csrc/ag_core.h as host+device functions; this program drives exactly those functions with plain loops
// (one "thread" per unit position / alignment / component, in the same order the kernels guarantee) so that the position-parallel
// formulation — touch fusion, first-compatible-on-final-table edges, component-parallel walk — is checked against the oracle
// and the reference on CPU.  The CUDA kernels themselves (sort, scans, pools, launches) are covered by the `-m gpu` tests.
#include "../../aligngraph_b200/csrc/ag_core.h"
#include "../../aligngraph_b200/csrc/ag_pipeline.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <unistd.h>

class AgEmul {
public:
    int k = 5, iv = 50, cov = 20;
    AgUnitInput in{};
    ag_reads rd{};
    // products
    std::vector<ag_alnp> alnp; std::vector<ag_fast> fast;
    std::vector<std::vector<u32>> tiles;
    std::vector<u32> pos_node, node_pos, node_sref, eovf_head, eovf_target, eovf_next, walk_next, parent;
    std::vector<ag_nodeb> nodeb;  // final order
    std::vector<ag_nodec> node_c; std::vector<ag_nodew> node_w; std::vector<ag_cm1> cm1;
    std::vector<unsigned char> pos_term; std::vector<u32> indeg, fnext, fprev, msuf, mnode; std::vector<ag_chain> chain; std::vector<ag_hrec> hrec; std::vector<ag_hdet> hdet; bool use_chains = false;
    std::vector<ag_nodeb> ovf_node; std::vector<u32> ovf_next; u32 ovf_count = 0; int err = 0;
    u32 n_nodes = 0;
    bool fallback_used = false;
    u32 n_live = 0, n_heads = 0, max_chain = 0, n_flagged = 0, n_flag2 = 0, n_tiles_total = 0, n_many = 0, max_items = 0;
    bool force_all_edges = getenv("AG_EMUL_ALL_EDGES") != nullptr;   // run the generic edge sweep on every tile (must change nothing)

    void set_reads(const AgReads& r) { rd.bases = r.bases.data(); rd.nmask = r.nmask.data(); rd.len = r.len.data(); rd.stride2 = r.stride2; rd.stridem = r.stridem; }
    std::vector<u32> own_cm_start; std::vector<ag_cm> own_cm;
    void load_unit(const AgUnitInput& i) {
        in = i;
        if (!in.cm_start) {   // the product derives the position-ordered contiMer table on the device (k_cm_count / k_cm_fill / k_cm_sort)
            ag_expand_contimers(in.threads, in.n_threads, in.chain_pos, in.n_cm, in.n_pos, own_cm_start, own_cm);
            in.cm_start = own_cm_start.data(); in.cm = own_cm.data();
        }
    }

    ag_cmtab cmt() const { ag_cmtab t; t.start = in.cm_start; t.cm = in.cm; return t; }
    // same single-candidate shortcut as the kernels' for_candidates_fast
    template <class F> void for_candidates_fast(u32 q, const ag_cm1& ca, u32 mate, F f) {
        ag_cm1 cb; cb.cid = cb.coff = AG_NONE;
        if (mate != AG_NONE) cb = cm1[mate];
        if (ca.cid != AG_CM_MANY && cb.cid != AG_CM_MANY && !force_generic) {
            ag_nodem c; c.cid = ca.cid; c.coff = ca.coff; c.cid0 = cb.cid; c.coff0 = cb.coff; c.moff = mate; f(c);
        } else ag_for_candidates(cmt(), q, mate, f);
    }
    bool force_generic = getenv("AG_EMUL_FORCE_GENERIC") != nullptr;

    void build() {   // like AgDevice::build: grow the overflow pool and redo the sweep when it runs out
        for (size_t cap = 1 << 18;; cap *= 4) { err = 0; if (build_once(cap)) return; if (cap > ((size_t)1 << 26)) throw AgHostError{"emul: overflow pool exhausted"}; }
    }
    bool build_once(size_t pool_cap) {
        const u32 nA = (u32)in.n_aln, n_ref = in.n_ref, n_pos = in.n_pos;
        alnp.resize(nA); fast.resize(nA);
        const u32 n_tiles = (n_ref + AG_TPOS - 1) / AG_TPOS;
        tiles.assign(n_tiles, {});
        ag_cmtab ct = cmt();
        cm1.resize((size_t)n_pos + 1);
        std::vector<u32> many_prefix((size_t)n_pos + 1, 0);
        for (u32 p = 0; p < n_pos; p++) { cm1[p] = ag_make_cm1(ct, p); many_prefix[p + 1] = many_prefix[p] + (cm1[p].cid == AG_CM_MANY ? 1u : 0u); }
        std::vector<u32> lin_prefix((size_t)n_pos + 1, 0);   // exclusive scan of the break flags (k_cm1 + scan on the device)
        for (u32 p = 0; p < n_pos; p++) lin_prefix[p + 1] = lin_prefix[p] + (p ? ag_cm1_break(cm1[p - 1], cm1[p]) : 0u);
        const bool no_linear = getenv("AG_EMUL_NO_LINEAR") != nullptr;
        for (u32 i = 0; i < nA; i++) {  // k_prep + k_keys + sort
            ag_prep_out o = ag_prep(in.aln[i], in.ext, rd.len[in.aln[i].pair], (u32)k);
            alnp[i] = o.p; fast[i] = ag_fast_prep(o.p, o.lo, o.span, i);
            if (!o.any) continue;
            if (o.lo + o.span >= n_ref) throw AgHostError{"BOWTIE ALIGNMENT ERROR"};
            ag_fast_classify(fast[i], o.p, many_prefix.data(), lin_prefix.data(), cm1.data());
            if (no_linear) fast[i].simple &= ~AG_FAST_LINEAR;
            u32 t0, t1; ag_tile_range(o.lo, o.lo + o.span, n_tiles, t0, t1);
            for (u32 t = t0; t <= t1; t++) tiles[t].push_back(i);
        }
        ovf_node.assign(pool_cap, ag_nodeb{}); ovf_next.assign(pool_cap, 0); ovf_count = 0;
        ag_ovfpool pool; pool.node = ovf_node.data(); pool.next = ovf_next.data(); pool.count = &ovf_count; pool.cap = (u32)ovf_node.size(); pool.err = &err;
        pos_node.assign((size_t)n_pos + 1, 0);
        nodeb.clear();
        std::vector<unsigned char> tile_flag(n_tiles, 0);
        // k_build: tile by tile, warp by warp; a warp's 32 lanes (31 owned positions + the halo) see every tile alignment together,
        // the successor item travels from lane + 1 to lane (the kernel's shuffle)
        for (u32 tile = 0; tile < n_tiles; tile++)
            for (u32 warp = 0; warp < 8; warp++) {
                const u32 wq0 = tile * AG_TPOS + warp * AG_WPOS;
                u32 slot_mem[32][AG_NF * AG_NODE_SCAP];
                ag_slots sv[32]; ag_plist pl[32]; ag_cm1 ca[32];
                for (u32 l = 0; l < 32; l++) {
                    sv[l].base = slot_mem[l]; sv[l].fstride = AG_NODE_SCAP; sv[l].nstride = 1;
                    pl[l].n = 0; pl[l].ovf_head = pl[l].ovf_tail = AG_NONE;
                    ca[l].cid = ca[l].coff = AG_NONE;
                    if (wq0 + l < n_ref) ca[l] = cm1[wq0 + l];
                }
                for (u32 idx : tiles[tile]) {
                    const ag_fast f = fast[idx];
                    if (!(f.lo <= wq0 + 31 && f.lo + f.span >= wq0)) continue;
                    u32 item[32]; bool want[32];
                    for (u32 l = 0; l < 32; l++) {
                        const u32 q = wq0 + l;
                        item[l] = AG_NONE; want[l] = false;
                        if (q < n_ref && q - f.lo <= f.span) {
                            auto codef = [&](u32 soff) -> int { return rd.code(f.read, f.lsrc_len >> 16, soff); };
                            item[l] = ag_lane_touch(want[l], pl[l], sv[l], pool, ct, cm1.data(), ca[l], f, &alnp[idx], in.ext, q, (u32)k, iv, force_generic, codef);
                        }
                    }
                    for (u32 l = 0; l < AG_WPOS; l++) {
                        if (!want[l]) continue;
                        const u32 nb = item[l + 1];
                        if (item[l] != AG_NONE && nb < 32u) ag_note_succ(pl[l], sv[l], pool, item[l], nb);
                        else tile_flag[tile] |= (item[l] == AG_NONE) ? 1 : 2;
                    }
                }
                if (err) return false;
                for (u32 l = 0; l < AG_WPOS; l++) {  // write-out in position order (the kernel writes tile blocks in arbitrary order; k_succ moves them into this order)
                    const u32 q = wq0 + l;
                    if (q >= n_ref) break;
                    pos_node[q] = (u32)nodeb.size();
                    if (ca[l].cid != AG_CM_MANY) {
                        const u32 nloc = pl[l].n < (u32)AG_NODE_SCAP ? pl[l].n : (u32)AG_NODE_SCAP;
                        for (u32 i = 0; i < nloc; i++) {
                            ag_nodeb b; b.cid = ca[l].cid; b.coff = ca[l].coff; b.cid0 = sv[l].ld(AG_F_CID0, i); b.coff0 = sv[l].ld(AG_F_COFF0, i); b.moff = sv[l].ld(AG_F_MOFF, i);
                            b.cov = sv[l].ld(AG_F_COV, i);
                            for (u32 j = 0; j < 5; j++) b.cnt[j] = sv[l].ld(AG_F_CNT + j, i);
                            b.sread = sv[l].ld(AG_F_SREAD, i); b.soff_len = sv[l].ld(AG_F_SL, i); b.succ = sv[l].ld(AG_F_SUCC, i);
                            nodeb.push_back(b);
                        }
                    }
                    for (u32 o = pl[l].ovf_head; o != AG_NONE; o = ovf_next[o]) nodeb.push_back(ovf_node[o]);
                    if ((u32)nodeb.size() - pos_node[q] != pl[l].n) throw AgHostError{"emul: node list length mismatch"};
                }
            }
        for (u32 q = n_ref; q <= n_pos; q++) pos_node[q] = (u32)nodeb.size();
        n_nodes = (u32)nodeb.size();
        n_many = many_prefix[n_pos]; max_items = 0; n_flag2 = 0;
        for (u32 q = 0; q < n_ref; q++) max_items = std::max(max_items, pos_node[q + 1] - pos_node[q]);
        for (u32 t = 0; t < n_tiles; t++) if (tile_flag[t] & 2) n_flag2++;
        // emit_node
        node_c.resize(n_nodes); node_w.resize(n_nodes); node_sref.resize(2 * (size_t)n_nodes); node_pos.resize(n_nodes);
        eovf_head.assign(n_nodes, AG_NONE); eovf_target.clear(); eovf_next.clear();
        for (u32 q = 0; q < n_ref; q++)
            for (u32 v = pos_node[q]; v < pos_node[q + 1]; v++) {
                const ag_nodeb& b = nodeb[v];
                ag_nodec c; c.cid = b.cid; c.coff = b.coff; c.cid0 = b.cid0; c.coff0 = b.coff0; node_c[v] = c;
                ag_nodew w; w.succ0 = w.succ1 = AG_NONE; w.moff = b.moff;
                w.misc = ag_node_misc(b.cid, b.coff, b.cov, b.cnt, in.ref[q], cov); node_w[v] = w;
                node_sref[2 * (size_t)v] = b.sread; node_sref[2 * (size_t)v + 1] = b.soff_len; node_pos[v] = q;
            }
        // k_succ
        for (u32 v = 0; v < n_nodes; v++) {
            u32 mask = nodeb[v].succ;
            if (!mask) continue;
            const u32 b1 = pos_node[node_pos[v] + 1];
            for (u32 j = 0; j < 32; j++) if ((mask >> j) & 1) {
                if (b1 + j >= pos_node[node_pos[v] + 2]) throw AgHostError{"emul: successor item out of range"};
                if (ag_edge_ok_c(node_c[v], node_c[b1 + j], iv)) add_edge(v, b1 + j);
            }
        }
        // k_edges on the flagged tiles
        n_flagged = 0;
        for (u32 tile = 0; tile < n_tiles; tile++) {
            if (!tile_flag[tile] && !force_all_edges) continue;
            n_flagged++;
            const bool all = (tile_flag[tile] & 2) || force_all_edges;
            for (u32 q = tile * AG_TPOS; q < (tile + 1) * AG_TPOS && q < n_ref; q++) {
                u32 nb0 = pos_node[q], nn0 = pos_node[q + 1] - nb0;
                if (!nn0) continue;
                const ag_cm1 ca = cm1[q];
                u32 last_v = AG_NONE, last_t = AG_NONE;
                for (u32 idx : tiles[tile]) {
                    const ag_fast f = fast[idx];
                    if (!all && (f.simple & AG_FAST_CLEAN) && !force_generic) continue;   // settled inside the node sweep
                    bool simple = f.simple && !force_generic;
                    if (q - f.lo >= f.span + (simple ? 0u : 1u)) continue;
                    ag_touch t;
                    if (simple) t = ag_fast_touch(f, q, (u32)k);
                    else { t = ag_locate(alnp[idx], in.ext, q, (u32)k); if (t.kind != 1) continue; }
                    u32 nb1 = pos_node[t.npos], nn1 = pos_node[t.npos + 1] - nb1;
                    const ag_cm1 cn1 = cm1[t.npos];
                    for_candidates_fast(q, ca, t.mate, [&](const ag_nodem& c) {
                        u32 ci = ag_first_compatible(node_c.data(), node_w.data(), nb0, nn0, c, iv);
                        if (ci == AG_NONE) return;
                        const ag_nodec x = node_c[nb0 + ci];
                        for_candidates_fast(t.npos, cn1, t.nmate, [&](const ag_nodem& c2) {
                            u32 ni = ag_first_compatible(node_c.data(), node_w.data(), nb1, nn1, c2, iv);
                            if (ni == AG_NONE) return;
                            if (nb0 + ci == last_v && nb1 + ni == last_t) return;
                            if (ag_edge_ok_c(x, node_c[nb1 + ni], iv)) add_edge(nb0 + ci, nb1 + ni);
                            last_v = nb0 + ci; last_t = nb1 + ni;
                        });
                    });
                }
            }
        }
        n_tiles_total = n_tiles;
        return true;
    }

    void add_edge(u32 v, u32 tgt) {
        ag_nodew& w = node_w[v];
        if (w.succ0 == tgt || w.succ1 == tgt) return;
        if (w.succ0 == AG_NONE) { w.succ0 = tgt; return; }
        if (w.succ1 == AG_NONE) { w.succ1 = tgt; return; }
        if (w.misc & AG_NW_OVF) for (u32 o = eovf_head[v]; o != AG_NONE; o = eovf_next[o]) if (eovf_target[o] == tgt) return;
        eovf_target.push_back(tgt); eovf_next.push_back((w.misc & AG_NW_OVF) ? eovf_head[v] : AG_NONE);
        eovf_head[v] = (u32)eovf_target.size() - 1; w.misc |= AG_NW_OVF;
    }

    ag_walkctx ctx() {
        ag_walkctx w; w.nw = node_w.data(); w.node_pos = node_pos.data(); w.pos_node = pos_node.data(); w.ovf_head = eovf_head.data();
        w.ovf_target = eovf_target.data(); w.ovf_next = eovf_next.data(); w.cmt = cmt(); w.chain_pos = in.chain_pos; w.walk_next = walk_next.data(); w.chain = chain.data(); w.hrec = hrec.data(); w.hdet = hdet.data(); w.msuf = msuf.data(); w.mnode = mnode.data(); w.fprev = fprev.data();
        return w;
    }
    u32 find(u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
    void unite(u32 a, u32 b) { a = find(a); b = find(b); if (a == b) return; if (a > b) std::swap(a, b); parent[b] = a; }
    bool live(u32 v) const { return !(node_w[v].misc & AG_NW_FILTERED); }
    void fill_tail(ag_walk& r) { r.tail_sread = node_sref[2 * (size_t)r.last_node]; r.tail_soff_len = node_sref[2 * (size_t)r.last_node + 1]; }

    void extend(std::vector<ag_walk>& walks) {
        walks.clear();
        walk_next.assign(n_nodes, AG_NONE); parent.resize(n_nodes);
        std::iota(parent.begin(), parent.end(), 0u);
        ag_cmtab ct = cmt();
        // forced-link chains (k_indeg, k_links, list ranking)
        pos_term.assign((size_t)in.n_pos + 1, 0);
        for (u32 p = 0; p < in.n_pos; p++) for (u32 e = ct.start[p]; e < ct.start[p + 1]; e++) if (ct.cm[e].chain == ct.cm[e].term) pos_term[p] = 1;
        indeg.assign(n_nodes, 0); fnext.assign(n_nodes, AG_NONE); fprev.assign(n_nodes, AG_NONE); msuf.assign(n_nodes, 0); mnode.assign(n_nodes, AG_NONE); chain.assign(n_nodes, ag_chain{});
        for (u32 v = 0; v < n_nodes; v++) {
            if (!live(v)) continue;
            const ag_nodew& x = node_w[v];
            if (x.succ0 != AG_NONE && live(x.succ0)) indeg[x.succ0]++;
            if (x.succ1 != AG_NONE && live(x.succ1)) indeg[x.succ1]++;
            if (x.misc & AG_NW_OVF) for (u32 o = eovf_head[v]; o != AG_NONE; o = eovf_next[o]) if (live(eovf_target[o])) indeg[eovf_target[o]]++;
        }
        for (u32 v = 0; v < n_nodes; v++) {
            fnext[v] = ag_forced_succ(node_w.data(), eovf_head.data(), eovf_target.data(), eovf_next.data(), indeg.data(), pos_term.data(), node_pos.data(), v);
            if (fnext[v] != AG_NONE) { node_w[fnext[v]].misc |= AG_NW_INTERIOR; fprev[fnext[v]] = v; }
        }
        for (u32 v = n_nodes; v-- > 0;) {  // forced links point to higher node indices (positions increase), so one backward pass ranks every chain
            ag_chain c; c.jump = AG_NONE; c.tail = v; c.len = 1; c.flg = (node_w[v].misc & AG_NW_HASCONTIG) ? 1u : 0u;
            if (fnext[v] != AG_NONE) {
                if (fnext[v] <= v) throw AgHostError{"emul: forced link does not point forward"};
                const ag_chain& j = chain[fnext[v]]; c.tail = j.tail; c.len += j.len; c.flg += j.flg;
            }
            chain[v] = c;
        }
        use_chains = true;
        // start candidates = chain heads in node order (k_cand_*)
        std::vector<u32> cand;
        n_live = 0; max_chain = 0;
        for (u32 v = 0; v < n_nodes; v++) if (live(v)) { n_live++; if (!(node_w[v].misc & AG_NW_INTERIOR)) { cand.push_back(v); max_chain = std::max(max_chain, chain[v].len); } }
        n_heads = (u32)cand.size();
        hrec.assign(n_nodes, ag_hrec{});   // k_hrec
        hdet.assign(n_nodes, ag_hdet{});
        for (u32 h : cand) {
            hrec[h] = ag_make_hrec(chain[h], node_w[chain[h].tail], ct, node_pos[chain[h].tail]);
            if (hrec[h].tcm != AG_NONE) hdet[h] = ag_make_hdet(ct, in.chain_pos, pos_node.data(), hrec[h].tcm);
        }
        // components over chain tails (k_uf_tails, k_uf_flatten)
        for (u32 h : cand) {
            const u32 t = chain[h].tail;
            const ag_nodew& x = node_w[t];
            if (x.succ0 != AG_NONE && live(x.succ0)) unite(t, chain[x.succ0].tail);
            if (x.succ1 != AG_NONE && live(x.succ1)) unite(t, chain[x.succ1].tail);
            if (x.misc & AG_NW_OVF) for (u32 o = eovf_head[t]; o != AG_NONE; o = eovf_next[o]) if (live(eovf_target[o])) unite(t, chain[eovf_target[o]].tail);
            u32 p = node_pos[t], c0 = ct.start[p];
            if (ct.start[p + 1] - c0 != 1) continue;
            ag_cm m = ct.cm[c0];
            if (m.chain == m.term) continue;
            u32 z = in.chain_pos[m.term];
            for (u32 y = pos_node[z]; y < pos_node[z + 1]; y++) if (live(y)) unite(t, chain[y].tail);
        }
        std::vector<u32> label(cand.size()), cmin(n_nodes, AG_NONE), cmax(n_nodes, 0);
        for (u32 i = 0; i < cand.size(); i++) { u32 r = find(chain[cand[i]].tail); label[i] = r; cmin[r] = std::min(cmin[r], i); cmax[r] = std::max(cmax[r], i); }
        ag_walkctx w = ctx();
        // components are replayed from the LAST candidate's root to the first to make sure nothing depends on cross-component order
        if (getenv("AG_EMUL_COMPONENT_STATS")) {
            std::vector<u32> sz(n_nodes, 0); u32 ncomp = 0, mx = 0, mxspan = 0; std::vector<u32> hist(8, 0);
            for (u32 i = 0; i < cand.size(); i++) sz[label[i]]++;
            for (u32 v = 0; v < n_nodes; v++) if (sz[v]) { ncomp++; mx = std::max(mx, sz[v]); mxspan = std::max(mxspan, cmax[v] - cmin[v] + 1); u32 b = 0; while ((1u << (2 * b + 2)) <= sz[v] && b < 7) b++; hist[b]++; }
            fprintf(stderr, "components: %u over %zu candidates, largest %u candidates, widest candidate range %u; sizes <4,<16,<64,<256,<1k,<4k,<16k,more:", ncomp, cand.size(), mx, mxspan);
            for (u32 b = 0; b < 8; b++) fprintf(stderr, " %u", hist[b]);
            fprintf(stderr, "\n");
        }
        for (u32 i0 = (u32)cand.size(); i0-- > 0;) {
            const u32 r = chain[cand[i0]].tail;
            if (find(r) != r) continue;
            for (u32 i = cmin[r]; i <= cmax[r]; i++) {
                if (label[i] != r) continue;
                u32 v = cand[i];
                if (node_w[v].misc & AG_NW_TRAV) continue;
                ag_walk x = ag_walk_from(w, v); fill_tail(x); walks.push_back(x);
            }
        }
        std::sort(walks.begin(), walks.end(), [](const ag_walk& a, const ag_walk& b) { return a.start_node < b.start_node; });
        // skip-rule trigger (AG:2194-2202)
        u32 bso = 0, beo = 0; bool have = false, trigger = false;
        for (const ag_walk& r : walks) {
            u32 eoff = r.eoff;
            if (((r.flags >> 1) & 3) != 1) eoff = eoff + (r.tail_soff_len >> 16) - 1;
            if (have && bso <= r.soff && beo >= eoff) continue;
            bso = r.soff; beo = eoff; have = true;
            if (beo - bso > 100000u) { trigger = true; break; }
        }
        if (trigger || getenv("AG_EMUL_FORCE_SEQUENTIAL")) {
            fallback_used = true; use_chains = false;
            walks.clear();
            for (u32 v = 0; v < n_nodes; v++) { u32 m = node_w[v].misc & ~(AG_NW_TRAV | AG_NW_DETOUR | AG_NW_STOP); node_w[v].misc = live(v) ? m : (m | AG_NW_TRAV); walk_next[v] = AG_NONE; msuf[v] = 0; mnode[v] = AG_NONE; }
            w = ctx();
            u32 sbo = AG_NONE, seo = AG_NONE, sei = AG_NONE;
            for (u32 cp = 0; cp < in.n_ref;) {   // k_walk_sequential
                for (u32 v = pos_node[cp]; v < pos_node[cp + 1]; v++) {
                    if (ag_seq_trav(w, v)) continue;
                    ag_walk x = ag_walk_from_seq(w, v); fill_tail(x); walks.push_back(x);
                    u32 eoff = x.eoff;
                    if (((x.flags >> 1) & 3) != 1) eoff = eoff + (x.tail_soff_len >> 16) - 1;
                    bool contained = (sei == 0) && sbo <= x.soff && seo >= eoff;
                    if (!contained) { sbo = x.soff; seo = eoff; sei = 0; }
                }
                if (seo - sbo > 100000u) { if (sei == 0 && cp + 1000 < seo) cp += 1000; else cp++; }
                else cp++;
            }
        }
    }

    std::string bases_buf;
    void materialize(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char*& bases_out, std::vector<u64>& offs) {
        offs.assign(sel.size() + 1, 0);
        for (size_t i = 0; i < sel.size(); i++) offs[i + 1] = offs[i] + walks[sel[i]].len + ag_walk_tail_len(walks[sel[i]]);
        std::string& bases = bases_buf;
        bases.assign(offs.back(), '\0');
        bases_out = bases.empty() ? nullptr : &bases[0];
        ag_cmtab ct = cmt();
        std::vector<u32> tail_end(n_nodes, AG_NONE);
        for (size_t i = 0; i < sel.size(); i++) {
            size_t o = offs[i];
            if (use_chains) {   // k_mat_items: hop from chain to chain, note where every chain's run of bases ends; detours copied right away
                for (u32 v = walks[sel[i]].start_node; v != AG_NONE;) {
                    const ag_chain c = chain[v];
                    o += c.len;
                    const u32 t = c.tail;
                    if (tail_end[t] != AG_NONE) throw AgHostError{"emul: a chain was emitted twice"};
                    tail_end[t] = (u32)o;
                    if (node_w[t].misc & AG_NW_DETOUR) { ag_cm m = ct.cm[ct.start[node_pos[t]]]; for (u32 e = m.chain + 1; e <= m.term; e++) bases[o++] = in.chain_base[e]; }
                    v = walk_next[t];
                }
            } else {            // k_materialize_seq
                for (u32 v = walks[sel[i]].start_node; v != AG_NONE;) {
                    bases[o++] = (char)(node_w[v].misc & 0xFF);
                    if (!(node_w[v].misc & AG_NW_STOP)) { v = fnext[v]; continue; }
                    if (node_w[v].misc & AG_NW_DETOUR) { ag_cm m = ct.cm[ct.start[node_pos[v]]]; for (u32 e = m.chain + 1; e <= m.term; e++) bases[o++] = in.chain_base[e]; }
                    v = walk_next[v];
                }
            }
            {   // k_mat_tails
                const ag_walk& r = walks[sel[i]];
                u32 tl = ag_walk_tail_len(r), slen = r.tail_soff_len >> 16, soff = r.tail_soff_len & 0xFFFFu;
                if (tl) { u32 rlen = rd.len[r.tail_sread >> 2]; for (u32 j = 1; j < slen; j++) bases[o++] = "ACGTN"[rd.code(r.tail_sread, rlen, soff + j)]; }
            }
            if (o != offs[i + 1]) throw AgHostError{"emul: walk length mismatch"};
        }
        if (use_chains)   // k_mat_nodes: every node of an emitted chain writes its own byte
            for (u32 v = 0; v < n_nodes; v++) { const u32 e = tail_end[chain[v].tail]; if (e != AG_NONE) bases[e - chain[v].len] = (char)(node_w[v].misc & 0xFF); }
        for (size_t i = 0; i < bases.size(); i++) if (!bases[i]) throw AgHostError{"emul: hole in the materialised bases"};
    }

    void materialize_begin(const std::vector<ag_walk>& walks, const std::vector<u32>& sel, char*& bases_out, std::vector<u64>& offs) { materialize(walks, sel, bases_out, offs); }
    void materialize_wait() {}
    void occupancy_begin() {}
    void occupancy_wait(std::vector<unsigned char>& bits) { occupancy(bits); }
    void occupancy(std::vector<unsigned char>& bits) {
        bits.assign(((size_t)in.n_pos + 7) / 8, 0);
        for (u32 p = 0; p < in.n_pos; p++) {
            bool occ = in.cm_start[p + 1] > in.cm_start[p];
            if (p < in.n_ref && pos_node[p + 1] > pos_node[p]) occ = true;
            if (occ) bits[p >> 3] |= (unsigned char)(1u << (p & 7));
        }
    }

    void dump_nodes(AgNodeDump& d) {
        d = AgNodeDump(); d.edge_start.push_back(0);
        for (u32 v = 0; v < n_nodes; v++) {
            const ag_nodeb& b = nodeb[v];
            d.pos.push_back(node_pos[v]); d.item.push_back(v - pos_node[node_pos[v]]); d.cov.push_back(b.cov);
            for (int j = 0; j < 5; j++) d.cnt.push_back(b.cnt[j]);
            d.cid.push_back(b.cid); d.coff.push_back(b.coff); d.cid0.push_back(b.cid0); d.coff0.push_back(b.coff0); d.moff.push_back(b.moff);
            d.sread.push_back(b.sread); d.soff_len.push_back(b.soff_len);
            std::vector<u32> e;
            if (node_w[v].succ0 != AG_NONE) e.push_back(node_w[v].succ0);
            if (node_w[v].succ1 != AG_NONE) e.push_back(node_w[v].succ1);
            if (node_w[v].misc & AG_NW_OVF) for (u32 o = eovf_head[v]; o != AG_NONE; o = eovf_next[o]) e.push_back(eovf_target[o]);
            std::sort(e.begin(), e.end());
            for (u32 x : e) d.edge_target.push_back(x);
            d.edge_start.push_back((u32)d.edge_target.size());
        }
    }
};

int main(int argc, char** argv) {
    std::string dir = ".";
    int dump = 0, first = 0, last = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--dir") dir = argv[++i];
        else if (a == "--dump-nodes") dump = 1;
        else if (a == "--check-tiles") {   // ag_tile_range: exactly the tiles whose owned positions or halo position meet the touched range
            unsigned long long x = 0x2545F4914F6CDD1Dull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (unsigned)(x >> 33); };
            size_t bad = 0, total = 2000000;
            for (size_t it = 0; it < total; it++) {
                const u32 n_ref = 1 + rnd() % 5000, n_tiles = (n_ref + AG_TPOS - 1) / AG_TPOS;
                const u32 lo = rnd() % n_ref, hi = lo + rnd() % (n_ref - lo);
                u32 t0, t1; ag_tile_range(lo, hi, n_tiles, t0, t1);
                for (u32 t = 0; t < n_tiles; t++) {
                    const u32 a0 = t * AG_TPOS, a1 = a0 + AG_TPOS;                 // owned [a0, a1) + halo a1
                    const bool meets = lo <= a1 && hi >= a0;
                    if (meets != (t >= t0 && t <= t1)) bad++;
                }
            }
            printf("%s checked=%zu bad=%zu\n", bad ? "DIFFERENT" : "IDENTICAL", total, bad);
            return bad ? 1 : 0;
        }
        else if (a == "--check-clean") {   // ag_fast_is_clean (prefix counts over both mates' ranges) against its definition, position by position
            unsigned long long x = 0xD1B54A32D192ED03ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (unsigned)(x >> 33); };
            const u32 n_pos = 5000; size_t bad = 0, total = 300000, n_clean = 0;
            for (size_t it = 0; it < total; it++) {
                std::vector<u32> many(n_pos + 1, 0), pre(n_pos + 2, 0);
                const unsigned density = rnd() % 4 == 0 ? 0 : 1 + rnd() % 40;
                for (u32 p = 0; p < n_pos; p++) many[p] = density && rnd() % (density * 25) == 0;
                for (u32 p = 0; p <= n_pos; p++) pre[p + 1] = pre[p] + many[p];
                const u32 len = 30 + rnd() % 200, k = 1 + rnd() % 9;
                ag_aln al{}; al.pair = 0;
                const u32 c5a = rnd() % 3 ? 0 : rnd() % 12, c3a = rnd() % 3 ? 0 : rnd() % 12, c5b = rnd() % 3 ? 0 : rnd() % 12, c3b = rnd() % 3 ? 0 : rnd() % 12;
                const u32 la = len - c5a - c3a, lb = len - c5b - c3b;
                al.dst1 = 300 + rnd() % 3000; al.sl1 = c5a | (la << 16);
                al.dst2 = al.dst1 + (rnd() % 3 ? rnd() % 900 : 0) - (rnd() % 5 == 0 ? rnd() % 250 : 0); al.sl2 = c5b | (lb << 16);
                al.flags = (rnd() & 1) | (1u << 8) | (1u << 16);
                const ag_prep_out o = ag_prep(al, nullptr, len, k);
                if (!o.any) continue;
                ag_fast f = ag_fast_prep(o.p, o.lo, o.span);
                bool def = f.simple != 0;
                for (u32 q = f.lo; q <= f.lo + f.span && def; q++) { if (many[q]) def = false; const u32 m = ag_fast_mate(f, q); if (m != AG_NONE && many[m]) def = false; }
                const bool got = ag_fast_is_clean(f, o.p, pre.data());
                n_clean += got;
                if (got != def) { if (bad < 5) fprintf(stderr, "differs: lo %u span %u mlo %u mlen %u mdelta %u def %d got %d\n", f.lo, f.span, f.mlo, f.mlen, f.mdelta, (int)def, (int)got); bad++; }
            }
            printf("%s checked=%zu clean=%zu bad=%zu\n", bad ? "DIFFERENT" : "IDENTICAL", total, n_clean, bad);
            return bad ? 1 : 0;
        }
        else if (a == "--check-sam-lines") {   // fuzz: random SAM records (CIGAR soup, odd RNAMEs, missing fields) through both record parsers
            unsigned long long x = 0x9E3779B97F4A7C15ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (unsigned)(x >> 33); };
            size_t bad = 0, total = 200000;
            for (size_t it = 0; it < total; it++) {
                std::string l = std::to_string(rnd() % 3000000) + "\t" + std::to_string(rnd() % 4096) + "\t";
                switch (rnd() % 6) { case 0: l += "*"; break; case 1: l += "0"; break; case 2: l += "7.3"; break; case 3: l += "chr*1"; break; case 4: l += "12"; break; default: l += ".5"; }
                l += "\t" + std::to_string(rnd() % 5000000) + "\t44\t";
                const unsigned nops = rnd() % 9 == 0 ? 30 + rnd() % 40 : rnd() % 8;
                for (unsigned k = 0; k < nops; k++) { if (rnd() % 11) l += std::to_string(rnd() % 160); l += "MMMMIDSS*MX="[rnd() % (rnd() % 50 ? 10 : 12)]; }
                if (rnd() % 7) l += "\t=\t100\t300\t*\t*\tAS:i:0";
                if (rnd() % 97 == 0) l = l.substr(0, rnd() % (l.size() + 1));   // truncated record
                if (!ag_selfcheck_sam_line(l.data(), l.size())) { if (bad < 5) fprintf(stderr, "differs: %s\n", l.c_str()); bad++; }
            }
            printf("%s lines=%zu bad=%zu\n", bad ? "DIFFERENT" : "IDENTICAL", total, bad);
            return bad ? 1 : 0;
        }
        else if (a == "--check-code4") {   // the eight-bases-at-a-time oriented 4-bit coder against its per-base definition and against ag_reads::code
            std::vector<std::string> seqs; unsigned long long x = 88172645463325252ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (unsigned)(x >> 33); };
            for (int i = 0; i < 4000; i++) { size_t n = 1 + rnd() % 256; std::string s1(n, 'A'), s2(n, 'A'); for (auto& ch : s1) ch = "ACGTACGTACGTNacgR"[rnd() % 17]; for (auto& ch : s2) ch = "ACGTACGTACGTNacgR"[rnd() % 17]; seqs.push_back(s1); seqs.push_back(s2); }
            AgReads q; ag_pack_reads(seqs, q);
            ag_reads rd; rd.bases = q.bases.data(); rd.nmask = q.nmask.data(); rd.len = q.len.data(); rd.stride2 = q.stride2; rd.stridem = q.stridem;
            size_t bad = 0, checked = 0;
            for (u32 read = 0; read < 2 * q.n_pairs; read++) {
                const u32 len = q.len[read >> 1]; const u32* b = rd.bases + (size_t)read * rd.stride2; const u32* m = rd.nmask + (size_t)read * rd.stridem;
                for (u32 rc = 0; rc < 2; rc++)
                    for (u32 j = 0; j < (len + 7) / 8; j++) {
                        const u32 w = ag_code4_word(b, m, rd.stride2, rd.stridem, rc, len, j), wr = ag_code4_word_ref(b, m, rc, len, j);
                        for (u32 t = 0; t < 8 && 8 * j + t < len; t++) {
                            const u32 c = (w >> (4 * t)) & 7, cr = (wr >> (4 * t)) & 7, cd = (u32)rd.code((read << 1) | rc, len, 8 * j + t);
                            checked++; if (c != cr || c != cd) bad++;
                        }
                    }
            }
            printf("%s checked=%zu bad=%zu\n", bad ? "DIFFERENT" : "IDENTICAL", checked, bad);
            return bad ? 1 : 0;
        }
        else if (a == "--check-read-packers") {   // the multi-threaded word-at-a-time packer against the character-at-a-time one, on one FASTA file
            const std::string path = argv[++i];
            std::vector<std::string> seqs;
            { FILE* f = fopen(path.c_str(), "rb"); if (!f) return 2; char* line = nullptr; size_t cap = 0; ssize_t n;
              while ((n = getline(&line, &cap, f)) > 0) { if (line[n - 1] == '\n') n--; if (n && line[0] != '>') seqs.emplace_back(line, (size_t)n); } free(line); fclose(f); }
            setenv("AG_PARSE_PARALLEL_MIN", "0", 1);
            AgReads p, q; ag_parse_reads(path, p); ag_pack_reads(seqs, q);
            const bool ok = p.n_pairs == q.n_pairs && p.stride2 == q.stride2 && p.stridem == q.stridem && p.bases.size() == q.bases.size() &&
                            memcmp(p.bases.data(), q.bases.data(), p.bases.size() * 4) == 0 && memcmp(p.nmask.data(), q.nmask.data(), p.nmask.size() * 4) == 0 && p.len == q.len && p.exc == q.exc;
            printf("%s pairs=%lu exceptions=%zu\n", ok ? "IDENTICAL" : "DIFFERENT", (unsigned long)p.n_pairs, p.exc.size());
            return ok ? 0 : 1;
        }
        else if (a == "--contain-search") {   // <db.fa> <query.fa> <out.psl>: the built-in containment search of refinement(), candidates verified on the host
            const std::string db = argv[++i], q = argv[++i], out = argv[++i];
            try { ag_contain_search(db, q, out, ag_verify_placements_host, nullptr); } catch (const AgHostError& e) { printf("%s\n", e.msg.c_str()); return 255; }
            return 0;
        }
        else if (a == "--remove-misassembly") {   // <file> <id> <coverage> <run-aligners>: removeMisassembly's host logic in the current directory.
            // run-aligners = 0: on aligner outputs already in tmp/ (left there by the reference); 1: run the aligners on $PATH with the
            // reference's command lines (AlignGraph.cpp:3825-3849; the harness puts its stubs there), after writing tmp/_reads_1.fa,
            // tmp/_reads_2.fa and tmp/_genome.fa from reads_1.fa / reads_2.fa / genome.fa as a fresh run does
            const std::string file = argv[++i], id = argv[++i]; const int cov = atoi(argv[++i]); const int run = atoi(argv[++i]);
            try {
                if (run) {
                    if (system("mkdir -p tmp")) {}
                    std::vector<std::string> gids;
                    ag_formalize_reads("reads_1.fa", "reads_2.fa", "tmp");
                    ag_formalize_genome("genome.fa", "tmp", 1, gids);
                }
                ag_remove_misassembly(file, id, cov, "tmp", run ? [](const std::string& id_, void*) -> bool {
                    std::string c = "bowtie2-build -f tmp/_" + id_ + "_contigs.fa tmp/_" + id_ + "_contigs > bowtie_doc.txt 2> bowtie_doc.txt";
                    if (system(c.c_str())) {}
                    c = "bowtie2 -f --no-mixed -k 1 -p 8 -I 0 -X 1500 --no-discordant -x tmp/_" + id_ + "_contigs -1 tmp/_reads_1.fa -2 tmp/_reads_2.fa --reorder > tmp/_reads_" + id_ + "_contigs.bowtie 2> bowtie_doc.txt";
                    if (system(c.c_str())) {}
                    c = "pblat tmp/_genome.fa tmp/_" + id_ + "_contigs.fa -noHead tmp/_" + id_ + "_contigs_genome.psl -fastMap -threads=8 > blat_doc.txt 2> blat_doc.txt";
                    return system(c.c_str()) == 0;
                } : [](const std::string&, void*) -> bool { return true; }, ag_coverage_pileup_host, nullptr);
            }
            catch (const AgHostError& e) { printf("%s\n", e.msg.c_str()); return 255; }
            return 0;
        }
        else if (a == "--first") first = atoi(argv[++i]);
        else if (a == "--last") last = atoi(argv[++i]);
        else { fprintf(stderr, "ag_emul: unknown option %s\n", a.c_str()); return 2; }
    }
    ag_tune_malloc();
    try {
        if (chdir(dir.c_str()) != 0) throw AgHostError{"CANNOT OPEN FILE!"};
        int k = 5, iv = 50, cov = 20, part = 1;
        std::string contigFile, genomeFile;
        {
            FILE* f = fopen("tmp/_command.txt", "r");
            if (!f) throw AgHostError{"CANNOT OPEN FILE!"};
            char line[4096]; std::vector<std::string> tok;
            while (fgets(line, sizeof line, f)) { std::string s(line); while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); tok.push_back(s); }
            fclose(f);
            for (size_t i = 0; i + 1 < tok.size(); i++) {
                if (tok[i] == "--kMer") k = atoi(tok[i + 1].c_str());
                else if (tok[i] == "--insertVariation") iv = atoi(tok[i + 1].c_str());
                else if (tok[i] == "--coverage") cov = atoi(tok[i + 1].c_str());
                else if (tok[i] == "--part") part = atoi(tok[i + 1].c_str());
                else if (tok[i] == "--contig") contigFile = tok[i + 1];
                else if (tok[i] == "--genome") genomeFile = tok[i + 1];
            }
        }
        std::vector<std::string> cids, gids;
        ag_formalize_contigs(contigFile, "tmp", cids);
        int units = ag_formalize_genome(genomeFile, "tmp", part, gids);
        if (last < 0 || last >= units) last = units - 1;
        AgReads reads;
        ag_parse_reads("tmp/_reads.fa", reads);
        AgEmul eng; eng.k = k; eng.iv = iv; eng.cov = cov; eng.set_reads(reads);
        for (int unit = first; unit <= last; unit++) {
            AgUnitResult r;
            ag_run_unit_files(eng, reads, "tmp", unit, r);
            fprintf(stderr, "emul unit %d: aln=%lu nodes=%u live=%u chain_heads=%u max_chain=%u flagged_tiles=%u/%u flag2_tiles=%u many_pos=%u max_items=%u walks=%lu emitted=%lu%s\n", unit, (unsigned long)r.n_aln, eng.n_nodes, eng.n_live, eng.n_heads, eng.max_chain, eng.n_flagged, eng.n_tiles_total, eng.n_flag2, eng.n_many, eng.max_items, (unsigned long)r.n_walks, (unsigned long)r.n_emitted,
                    eng.fallback_used ? " (sequential walk)" : "");
            if (dump) { AgNodeDump d; eng.dump_nodes(d); std::string text; ag_format_node_dump(d, reads, text); ag_write_file("tmp/_nodes." + std::to_string(unit) + ".txt", text); }
        }
    } catch (const AgHostError& e) { printf("%s\n", e.msg.c_str()); return 255; }
    catch (const AgError& e) { printf("%s\n", e.msg.c_str()); return 255; }
    return 0;
}

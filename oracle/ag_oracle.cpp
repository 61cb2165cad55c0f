// ag_oracle — CPU restatement of AlignGraph's per-chromosome graph-build + extension path.
//
// *** TEST INFRASTRUCTURE.  NOT PRODUCT CODE. ***  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build or execute this program.  The product (aligngraph_b200/) never links it.
//
// What it restates (reference = /root/reference/AlignGraph/AlignGraph.cpp, cited as AG:line):
//   formalize contigs / genome            AG:3228-3345, AG:3347-3418   (what `--resume` re-runs, AG:4757-4758)
//   loadGenome                            AG:287-320
//   loadSeq / parseBLAT / loadContiAli    AG:322-359, AG:406-522, AG:817-852
//   keepPositions / updateContig          AG:731-748, AG:763-815
//   updateGenomeWithContig                AG:884-1217   (-> tmp/_initial_contigs.N.fa)
//   loadSeq(batch) / parseBOWTIE / loadReadAli / loadReadAlignment   AG:361-404, AG:181-285, AG:1233-1277, AG:1872-1895
//   compatible / updateKBases / updateKMer / updateGenomeWithRead    AG:1293-1312, AG:1340-1351, AG:1353-1624, AG:1635-1870
//   filterLowCoverage / max / contain / extdContigs1                 AG:1904-1918, AG:1944-1952, AG:1897-1902, AG:1954-2204
//   extdContigs2 (containment + join)     AG:2296-2380  (the file round trip AG:2206-2294 is an identity on u32 fields)
//   overlap / scaffoldContigs             AG:2388-2464  (-> tmp/_extended_contigs.N.fa)
//
// It is written as a *literal sequential* restatement (array-of-lists, same visiting order) so that every order-dependent
// behaviour of the reference is reproduced; the product uses a different, position-parallel formulation.
//
// Parity pin: tests/test_oracle.py runs the unmodified reference (oracle/_ref, built by `make ref`) and this
// program on the same tmp/ inputs and byte-compares all three per-unit outputs; tests/golden/ holds outputs of the real
// reference for the committed fixtures so the pin also holds where /root/reference is absent (the GPU box).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <chrono>
#include <unistd.h>

typedef uint32_t u32;
static const u32 NONE = 0xFFFFFFFFu;

// ---- tunables of the reference (AG:27-42) ---------------------------------------------------------------------------
static const double kInitContigThreshold = 0.5;  // INIT_CONTIG_THRESHOLD under OPTIMIZATION (AG:28-32)
static const double kContigThreshold = 0.5;      // CONTIG_THRESHOLD (AG:33)
static const double kReadThreshold = 0.6;        // THRESHOLD (AG:34)
static const int kBatch = 1000000;               // BATCH (AG:37)
static const int kEP = 5;                        // EP (AG:39)
static const long kLargeChunk = 1000000;         // LARGE_CHUNK (AG:40)

struct Seg { u32 src, dst, len; };
struct Pos { u32 id, off; };
struct Edge { u32 id, off, item; };
struct Node {
    u32 traversed = 0;
    std::string s;
    u32 cid = NONE, coff = NONE, cid0 = NONE, coff0 = NONE, mid = NONE, moff = NONE;
    std::vector<Edge> next;
    int cov = 0;
    int a = 0, c = 0, g = 0, t = 0, n = 0;
};
struct CMer { char base; u32 cid, coff, nid, noff, nitem; };
struct Slot { std::vector<Node> nodes; char base = 0; std::vector<CMer> cm; };
struct Seq { std::string bases; std::vector<std::vector<Pos>> sets; std::vector<int> fr; int outputted = 0; int id = 0; };
struct Contig { int extended; u32 sid, soff, eid, eoff, sid0, soff0, eid0, eoff0; std::string bases; };

static std::vector<Slot> G;          // the unit ("genome[0]" of the reference; every unit-local chromosome id is 0)
static std::vector<Contig> contigs;  // global `contigs` (AG:170)
static u32 lastSource = NONE;        // `sourceIDBak` (AG:762)
static std::string outdir = "tmp";

[[noreturn]] static void die(const char* msg) { std::cout << msg << std::endl; exit(255); }

static int absdiff(u32 a, u32 b) { return abs((int)(a - b)); }

// ---- input normalisation re-run by --resume ---------------------------------------------------------------------------
// AG:3228-3319 (contig branch): drop contigs <= 200 bp into _chaff.fa, cut the rest into <= 1 Mbp chunks named
// ">chunk.contig".
static void formalizeContigs(const std::string& path) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::vector<std::string> seqs, ids;
    std::string line;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '>') { seqs.emplace_back(); ids.push_back(line.substr(1)); }
        else seqs.back() += line;
    }
    std::ofstream out("tmp/_contigs.fa"), chaff("tmp/_chaff.fa");
    unsigned long chunk = 0, real = 0;
    for (size_t c = 0; c < seqs.size(); c++) {
        const std::string& s = seqs[c];
        if (s.size() > 200) {
            out << ">" << chunk++ << "." << real << "\n";
            if ((long)s.size() < kLargeChunk) {
                for (size_t i = 0; i < s.size(); i++) { out << s[i]; if ((i + 1) % 60 == 0 || i == s.size() - 1) out << "\n"; }
            } else {
                long total = 0;
                for (long i = 0; i < (long)s.size(); i++) {
                    out << s[i];
                    if ((i + 1) % kLargeChunk == 0 && i < (long)s.size() - 1 - 60) {
                        total += kLargeChunk;
                        out << "\n>" << chunk++ << "." << real << "\n";
                        continue;
                    }
                    if ((i + 1 - total) % 60 == 0 || i == (long)s.size() - 1) out << "\n";
                }
            }
            real++;
        } else {
            chaff << ">" << ids[c] << "\n";
            for (size_t i = 0; i < s.size(); i++) { chaff << s[i]; if ((i + 1) % 60 == 0 || i == s.size() - 1) chaff << "\n"; }
        }
    }
}

// AG:3347-3418: one file per unit, header ">0", chromosomes cut into `part` slices.
static int formalizeGenome(const std::string& path, int part) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::vector<std::string> chr;
    std::string line;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '>') chr.emplace_back(); else chr.back() += line;
    }
    int unit = 0;
    std::ofstream all("tmp/_genome.fa");
    for (size_t g = 0; g < chr.size(); g++) {
        const std::string& s = chr[g];
        std::ofstream out(("tmp/_genome." + std::to_string(unit) + ".fa").c_str());
        out << ">0\n"; all << ">" << unit << "\n";
        int q = 1;
        long slice = (long)s.size() / part;
        for (long i = 0; i < (long)s.size(); i++) {
            out << s[i]; all << s[i];
            bool cut = ((i + 1) % slice == 0 && q < part);
            if ((i + 1) % 60 == 0 || i == (long)s.size() - 1 || cut) { out << "\n"; all << "\n"; }
            if (i != (long)s.size() - 1 && cut) {
                out.close(); unit++; q++;
                out.open(("tmp/_genome." + std::to_string(unit) + ".fa").c_str());
                out << ">0\n"; all << ">" << unit << "\n";
            }
        }
        out.close(); unit++;
    }
    return unit;
}

// ---- loaders -------------------------------------------------------------------------------------------------------------
static void loadUnit(int unit) {  // AG:287-320
    std::ifstream in(("tmp/_genome." + std::to_string(unit) + ".fa").c_str());
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::string line;
    G.clear();
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '>') continue;
        for (char ch : line) { Slot s; s.base = ch; G.push_back(std::move(s)); }
    }
}

static void loadChunks(std::vector<Seq>& seqs) {  // AG:322-359
    std::ifstream in("tmp/_contigs.fa");
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::string line;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '>') {
            size_t dot = line.find('.');
            Seq s; s.id = atoi(line.substr(dot + 1).c_str());
            seqs.push_back(s);
        } else seqs.back().bases += line;
    }
}

// One PSL record (AG:406-522).  Fields are taken by tab index exactly as the reference does.
struct Psl { u32 tid, tstart, tend, tgap, sid, sstart, send, sgap, ssize, fr, tsize; std::vector<Seg> segs; };
static void parsePsl(const std::string& line, Psl& r) {
    std::vector<std::string> f;
    size_t a = 0;
    for (;;) { size_t b = line.find('\t', a); f.push_back(line.substr(a, b == std::string::npos ? b : b - a)); if (b == std::string::npos) break; a = b + 1; }
    auto fld = [&](size_t i) -> const char* { return i < f.size() ? f[i].c_str() : ""; };
    r.tid = atoi(fld(13)); r.tstart = atoi(fld(15)); r.tend = atoi(fld(16)); r.tgap = atoi(fld(7));
    r.sstart = atoi(fld(11)); r.send = atoi(fld(12)); r.sgap = atoi(fld(5)); r.ssize = atoi(fld(10)); r.tsize = atoi(fld(14));
    r.fr = NONE;
    if (f.size() > 8 && !f[8].empty()) r.fr = f[8][0] == '+' ? 0 : 1;
    std::string q = fld(9);
    size_t dot = q.find('.');
    r.sid = atoi(dot == std::string::npos ? q.c_str() : q.substr(0, dot).c_str());
    r.segs.clear();
    auto list = [&](size_t i, int which) {
        std::string s = fld(i); size_t p = 0, k = 0;
        for (;;) {
            size_t c = s.find(',', p);
            if (c == std::string::npos) break;
            u32 v = atoi(s.substr(p, c - p).c_str());
            if (which == 0) r.segs.push_back(Seg{NONE, NONE, v});
            else if (k < r.segs.size()) { if (which == 1) r.segs[k].src = v; else r.segs[k].dst = v; }
            k++; p = c + 1;
        }
    };
    list(18, 0); list(19, 1); list(20, 2);
}

static int keepSet(std::vector<Seq>& seqs, u32 sid, double thr) {  // AG:731-748
    if (sid == NONE) return 1;
    if (seqs[sid].sets.empty()) return 1;
    const std::vector<Pos>& last = seqs[sid].sets.back();
    int match = 0;
    for (const Pos& p : last) if (p.id != NONE) match++;
    return (double)match / last.size() >= thr ? 1 : 0;
}

static void addAlignment(std::vector<Seq>& seqs, u32 sid, u32 tid, const std::vector<Seg>& segs, u32 fr, double thr) {  // AG:763-815
    if (tid == NONE) return;
    auto open_set = [&]() {
        seqs[sid].sets.emplace_back(seqs[sid].bases.size(), Pos{NONE, NONE});
        seqs[sid].fr.push_back((int)fr);
    };
    if (sid != lastSource) {
        if (keepSet(seqs, lastSource, thr) == 0) { seqs[lastSource].sets.pop_back(); seqs[lastSource].fr.pop_back(); }
        open_set();
        lastSource = sid;
    } else {
        bool clash = false;
        for (size_t i = 0; i < segs.size() && !clash; i++)
            for (u32 j = segs[i].src; j < segs[i].src + segs[i].len; j++)
                if (seqs[sid].sets.back()[j].id != NONE) { clash = true; break; }
        if (clash) {
            if (keepSet(seqs, sid, thr) == 0) { seqs[sid].sets.pop_back(); seqs[sid].fr.pop_back(); }
            open_set();
        }
    }
    std::vector<Pos>& cur = seqs[sid].sets.back();
    for (const Seg& s : segs)
        for (u32 j = 0; j < s.len; j++) { cur[s.src + j].id = tid; cur[s.src + j].off = s.dst + j; }
}

static void loadContigAlignments(std::vector<Seq>& seqs, int unit) {  // AG:817-852
    std::ifstream in(("tmp/_contigs_genome." + std::to_string(unit) + ".psl").c_str());
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::string line;
    Psl r; r.sid = NONE;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) {
            if (keepSet(seqs, r.sid, kContigThreshold) == 0) seqs[r.sid].sets.pop_back();
            break;
        }
        parsePsl(line, r);
        if ((double)(r.send - r.sstart - r.sgap) / r.ssize >= kInitContigThreshold &&
            (double)(r.tend - r.tstart - r.tgap) / (r.tend - r.tstart) >= kInitContigThreshold && r.ssize > 200)
            addAlignment(seqs, r.sid, r.tid, r.segs, r.fr, kContigThreshold);
    }
}

static void revcomp(std::string& s) {  // AG:854-865
    std::reverse(s.begin(), s.end());
    for (char& c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

static void wrap60(std::ostream& o, const std::string& s) {
    for (size_t i = 0; i < s.size(); i++) { o << s[i]; if ((i + 1) % 60 == 0 || i == s.size() - 1) o << "\n"; }
}

// AG:884-1217.  SD = SI = 0 (AG:35-36) make every insertion "large" and every deletion "large".
static void threadContigs(std::vector<Seq>& seqs, int unit) {
    for (size_t sp = 0; sp < seqs.size(); sp++) {
        Seq& q = seqs[sp];
        for (size_t pp = 0; pp < q.sets.size(); pp++) {
            const std::vector<Pos>& ps = q.sets[pp];
            bool skip = false;
            for (size_t e = 0; e < pp && !skip; e++)
                if (absdiff(ps[0].off, q.sets[e][0].off) < (int)q.bases.size()) skip = true;  // AG:902-907
            for (size_t i = 0; !skip && i + 1 < ps.size(); i++)
                if (ps[i].id != NONE && G[ps[i].off].cm.size() >= 2) skip = true;          // AG:908-920
            if (skip) continue;
            bool flipped = false;
            if (q.fr[pp] == 1) { revcomp(q.bases); flipped = true; }
            q.outputted = 1;
            u32 curId = NONE, curOff = NONE, nid = NONE, noff = NONE;
            size_t i;
            for (i = 0; i + 1 < ps.size(); i++) {
                if (ps[i].id == NONE) continue;
                curId = ps[i].id; curOff = ps[i].off; nid = ps[i + 1].id; noff = ps[i + 1].off;
                char base = q.bases[i];
                if (nid == NONE) {                                                         // insertion to genome, AG:940-1045
                    for (size_t n = i + 2; n < ps.size(); n++) {
                        if (ps[n].id == NONE) continue;
                        nid = ps[n].id; noff = ps[n].off;
                        G[curOff].cm.push_back(CMer{base, (u32)sp, (u32)i, curId, (u32)G.size(), 0});
                        for (size_t j = 0; j < n - i - 2; j++) {
                            Slot b; b.base = q.bases[i + 1 + j];
                            b.cm.push_back(CMer{b.base, (u32)sp, (u32)(i + 1 + j), curId, (u32)G.size() + 1, 0});
                            G.push_back(std::move(b));
                        }
                        Slot b; b.base = q.bases[n - 1];
                        b.cm.push_back(CMer{b.base, (u32)sp, (u32)(n - 1), nid, noff, (u32)G[noff].cm.size()});
                        G.push_back(std::move(b));
                        i = n - 1;
                        break;
                    }
                } else {                                                                   // deletion / ordinary, AG:1046-1118
                    G[curOff].cm.push_back(CMer{base, (u32)sp, (u32)i, nid, noff, (u32)G[noff].cm.size()});
                }
            }
            if (nid != NONE) G[noff].cm.push_back(CMer{G[noff].base, (u32)sp, (u32)i, NONE, NONE, NONE});       // AG:1121-1134
            else G[curOff].cm.push_back(CMer{G[curOff].base, (u32)sp, (u32)i, NONE, NONE, NONE});               // AG:1135-1148
            if (flipped) revcomp(q.bases);
        }
    }
    // AG:1179-1216: original contigs of which >= 50 % of the chunks were threaded
    std::ofstream out((outdir + "/_initial_contigs." + std::to_string(unit) + ".fa").c_str());
    std::vector<std::string> whole; std::vector<int> placed, total;
    int prev = -1;
    for (Seq& q : seqs) {
        if (q.id != prev) { whole.emplace_back(); placed.push_back(0); total.push_back(0); prev = q.id; }
        total.back()++; placed.back() += q.outputted; whole.back() += q.bases;
    }
    for (size_t c = 0; c < whole.size(); c++)
        if ((double)placed[c] / (double)total[c] >= kContigThreshold) { out << ">" << c << "\n"; wrap60(out, whole[c]); }
}

// ---- read side -------------------------------------------------------------------------------------------------------------
struct Sam { u32 tid, tstart, tend, tgap, sid, sstart, send, sgap, ssize, fr; };
// AG:181-285.  `segs` is appended to, never cleared here (the caller clears, AG:1267-1268).
static void parseSam(const std::string& line, Sam& r, std::vector<Seg>& segs) {
    int item = 0, ins = 0, del = 0, total = 0, start = 0, end = 0, lead = 1, dotted = 0;
    std::string qname, flag, rname, pos, num;
    for (size_t i = 0; i < line.size(); i++) {
        char ch = line[i];
        if (ch == '\t') { item++; continue; }
        if (ch == '\0') break;
        if (item == 0) qname += ch;
        else if (item == 1) flag += ch;
        else if (item == 2) {
            if (ch == '*') {
                r.sid = atoi(qname.c_str());
                r.fr = (atoi(flag.c_str()) & 0x10) ? 1 : 0;
                r.tid = r.tstart = r.tend = r.tgap = r.sstart = r.send = r.sgap = r.ssize = NONE;
                return;
            }
            if (ch == '.') dotted = 1;
            if (!dotted) rname += ch;
        } else if (item == 3) pos += ch;
        else if (item == 5) {
            if (ch >= '0' && ch <= '9') num += ch;
            else if (ch == 'I') { ins += atoi(num.c_str()); total += atoi(num.c_str()); num.clear(); }
            else if (ch == 'D') { del += atoi(num.c_str()); num.clear(); }
            else if (ch == 'M') {
                Seg s; s.src = total; s.dst = atoi(pos.c_str()) + total + del - start - ins - 1; s.len = atoi(num.c_str());
                segs.push_back(s);
                total += atoi(num.c_str()); num.clear(); lead = 0;
            } else if (ch == 'S' && lead) { start = atoi(num.c_str()); total += start; num.clear(); lead = 0; }
            else if (ch == 'S') { end = atoi(num.c_str()); total += end; num.clear(); }
            else if (ch != '*') { std::cout << "unknown character: " << ch << std::endl; exit(255); }
        }
    }
    r.sid = atoi(qname.c_str());
    r.sstart = start; r.send = total - end; r.sgap = ins; r.ssize = total;
    bool star = !rname.empty() && rname[0] == '*';
    r.tid = star ? NONE : (dotted ? (u32)atoi(rname.c_str()) : 0);
    r.tstart = atoi(pos.c_str()) - 1;
    r.tend = r.tstart + total + del - ins;
    r.tgap = del;
    r.fr = (atoi(flag.c_str()) & 0x10) ? 1 : 0;
}

// AG:361-404: next batch of pairs from the interleaved read file; returns 0 when the batch is full.
static int loadReadBatch(std::ifstream& in, std::vector<Seq>& reads, int& firstId, int& lastId) {
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    firstId = lastId + 1;
    int second = 0;
    std::string line;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '>') reads.emplace_back();
        else {
            reads.back().bases += line;
            if (second) { lastId++; if ((lastId + 1) % kBatch == 0) return 0; second = 0; }
            else second = 1;
        }
    }
    return 1;
}

// AG:1233-1277
static int loadReadAlignments(std::ifstream& in, std::vector<Seq>& reads, int firstId, int lastId) {
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::vector<Seg> s1, s2;
    Sam a, b;
    std::string line;
    lastSource = NONE;
    while (in.good()) {
        std::getline(in, line);
        if (line[0] == 0) break;
        if (line[0] == '@') continue;
        parseSam(line, a, s1);
        std::getline(in, line);
        if (line[0] == 0) die("BROKEN BOWTIE FILE");
        parseSam(line, b, s2);
        if (a.sid < (u32)firstId) continue;
        if (a.sid > (u32)lastId) return 0;
        if (a.tid != NONE && b.tid != NONE &&
            (double)(a.send - a.sstart - a.sgap) / a.ssize >= kReadThreshold && (double)(a.tend - a.tstart - a.tgap) / (a.tend - a.tstart) >= kReadThreshold &&
            (double)(b.send - b.sstart - b.sgap) / b.ssize >= kReadThreshold && (double)(b.tend - b.tstart - b.tgap) / (b.tend - b.tstart) >= kReadThreshold) {
            addAlignment(reads, (a.sid - firstId) * 2, a.tid, s1, a.fr, kReadThreshold);
            addAlignment(reads, (b.sid - firstId) * 2 + 1, b.tid, s2, b.fr, kReadThreshold);
        }
        s1.clear(); s2.clear();
    }
    return 1;
}

static int compatible(const Node& x, const Node& y, int iv) {  // AG:1293-1312
    bool c1 = x.cid == NONE || y.cid == NONE || (x.cid == y.cid && absdiff(x.coff, y.coff) <= 5 * kEP) || (x.cid != y.cid);
    bool c2 = x.cid0 == NONE || y.cid0 == NONE || (x.cid0 == y.cid0 && absdiff(x.coff0, y.coff0) <= 2 * iv + 5 * kEP) || (x.cid0 != y.cid0);
    bool c3 = x.mid == NONE || y.mid == NONE || (x.mid == y.mid && absdiff(x.moff, y.moff) <= 2 * iv + 5 * kEP);
    return c1 && c2 && c3;
}

static void countBase(const std::string& s, Node& n) {  // AG:1340-1351
    if (s.empty()) return;
    switch (s[0]) { case 'A': n.a++; break; case 'C': n.c++; break; case 'G': n.g++; break; case 'T': n.t++; break; default: n.n++; }
}

static long g_events = 0, g_walks = 0, g_walk_bases = 0, g_walk_max = 0, g_chain_steps = 0, g_emitted = 0, g_emitted_bases = 0;

// One side of updateKMer (AG:1362-1477 for the node at P with bump=true, AG:1480-1587 for the node at nextP with
// bump=false): enumerate candidates = contiMers at P x contiMers at the mate position, first-compatible lookup, create
// when absent.  Returns the item indices.
static void touch(u32 off, u32 mid, u32 moff, const std::string& s, bool bump, int iv, std::vector<u32>& items) {
    Slot& slot = G[off];
    Node cand;
    cand.s = s; cand.mid = mid; cand.moff = moff; cand.cov = bump ? 1 : 0;
    size_t na = slot.cm.size();
    size_t nb = (mid != NONE) ? G[moff].cm.size() : 0;
    for (size_t ia = 0; ia < (na ? na : 1); ia++)
        for (size_t ib = 0; ib < (nb ? nb : 1); ib++) {
            cand.cid = na ? slot.cm[ia].cid : NONE; cand.coff = na ? slot.cm[ia].coff : NONE;
            cand.cid0 = nb ? G[moff].cm[ib].cid : NONE; cand.coff0 = nb ? G[moff].cm[ib].coff : NONE;
            size_t k = 0;
            for (; k < slot.nodes.size(); k++) if (compatible(cand, slot.nodes[k], iv)) break;
            if (k == slot.nodes.size()) { slot.nodes.push_back(cand); if (bump) countBase(s, slot.nodes[k]); }
            else if (bump) { slot.nodes[k].cov++; countBase(s, slot.nodes[k]); }
            items.push_back((u32)k);
        }
}

// AG:1353-1624
static void updateKMer(u32 off, u32 noff, u32 mid, u32 moff, u32 nmid, u32 nmoff, const std::string& s, const std::string& ns, int iv) {
    std::vector<u32> here, there;
    g_events++;
    touch(off, mid, moff, s, true, iv, here);
    touch(noff, nmid, nmoff, ns, false, iv, there);
    for (u32 ci : here)
        for (u32 ni : there) {
            Node& x = G[off].nodes[ci];
            const Node& y = G[noff].nodes[ni];
            bool present = false;
            for (const Edge& e : x.next) {
                if (e.id == NONE) die("KMER ERROR");
                if (e.id == 0 && e.off == noff && e.item == ni) { present = true; break; }
            }
            bool c1 = y.cid == NONE || x.cid == NONE || (y.cid == x.cid && absdiff(y.coff, x.coff) <= 5 * kEP) || (y.cid != x.cid);
            bool c2 = y.cid0 == NONE || x.cid0 == NONE || (y.cid0 == x.cid0 && absdiff(y.coff0, x.coff0) <= 2 * iv + 5 * kEP) || (y.cid0 != x.cid0);
            if (!present && c1 && c2) x.next.push_back(Edge{0, noff, ni});
        }
}

// AG:1635-1870
static void addReadBatch(std::vector<Seq>& reads, int k, int iv) {
    std::string s, ns;
    for (size_t sp = 0; sp + 1 < reads.size() || (sp < reads.size() && sp + 1 == reads.size()); sp += 2) {
        if (sp + 1 >= reads.size()) break;
        Seq& r1 = reads[sp]; Seq& r2 = reads[sp + 1];
        for (size_t pp = 0; pp < r1.sets.size(); pp++) {
            bool dup = false;
            for (size_t e = 0; e < pp && !dup; e++)
                if (absdiff(r1.sets[pp][0].off, r1.sets[e][0].off) < (int)r1.bases.size()) dup = true;  // AG:1650-1655
            if (dup) continue;
            int flip1 = 0, flip2 = 0, swapped = 0;
            if (r1.fr[pp] == 1 && r2.fr[pp] == 0) { revcomp(r1.bases); flip1 = 1; }
            else if (r2.fr[pp] == 1 && r1.fr[pp] == 0) { revcomp(r2.bases); flip2 = 1; }
            else die("BOWTIE ALIGNMENT ERROR");
            size_t span = r1.sets[pp].size() - (size_t)k;  // size_t arithmetic, as in the reference (AG:1672)
            for (size_t i = 0; i < span; i++)
                if (r1.sets[pp][i].id != NONE && r2.sets[pp][i].id != NONE && r1.sets[pp][i].off > r2.sets[pp][i].off) {
                    std::swap(r1.sets[pp], r2.sets[pp]); std::swap(r1.bases, r2.bases); swapped = 1; break;   // AG:1672-1679
                }
            const std::vector<Pos>& L = r1.sets[pp];
            const std::vector<Pos>& R = r2.sets[pp];
            const std::string& seq = r1.bases;
            for (size_t i = 0; i < span; i++) {
                if (L[i].id == NONE) continue;
                u32 off = L[i].off, mid = R[i].id, moff = R[i].off;
                u32 nid = L[i + 1].id, noff = L[i + 1].off, nmid = R[i + 1].id, nmoff = R[i + 1].off;
                if (nid == NONE) {                                                           // AG:1695-1790
                    for (size_t n = i + 2; n < L.size(); n++) {
                        if (L[n].id == NONE) continue;
                        noff = L[n].off; nmid = R[n].id; nmoff = R[n].off;
                        size_t stop = n + k < seq.size() ? n + k : seq.size();
                        s.assign(seq, i, k);
                        ns.assign(seq, n, stop - n);
                        if (noff == off + 1) updateKMer(off, noff, mid, moff, nmid, nmoff, s, ns, iv);   // AG:1707-1727
                        else {                                                                // AG:1730-1750
                            updateKMer(off, off + 1, mid, moff, NONE, NONE, s, std::string(), iv);
                            u32 c;
                            for (c = off + 1; c < noff - 1; c++) updateKMer(c, c + 1, NONE, NONE, NONE, NONE, std::string(), std::string(), iv);
                            updateKMer(c, c + 1, NONE, NONE, nmid, nmoff, std::string(), ns, iv);
                        }
                        i = n - 1;
                        break;
                    }
                } else {                                                                      // AG:1791-1857 (SD = 0)
                    s.assign(seq, i, k);
                    ns.assign(seq, i + 1, k);
                    updateKMer(off, noff, mid, moff, nmid, nmoff, s, ns, iv);
                }
            }
            if (swapped) { std::swap(r1.sets[pp], r2.sets[pp]); std::swap(r1.bases, r2.bases); }
            if (flip1 == 1 && flip2 == 0) revcomp(r1.bases);
            else if (flip2 == 1 && flip1 == 0) revcomp(r2.bases);
            else die("UNKNOWN ERROR");
        }
    }
}

static void loadReads(int unit, int k, int iv) {  // AG:1872-1895
    std::ifstream r("tmp/_reads.fa");
    std::ifstream ra(("tmp/_reads_genome." + std::to_string(unit) + ".bowtie").c_str());
    std::vector<Seq> reads;
    int lastId = -1, firstId = -1;
    for (;;) {
        int done = loadReadBatch(r, reads, firstId, lastId);
        loadReadAlignments(ra, reads, firstId, lastId);
        addReadBatch(reads, k, iv);
        if (done) break;
        reads.clear();
    }
}

// ---- extension -----------------------------------------------------------------------------------------------------------
static char consensus(const Node& n) {  // AG:1944-1952
    if (!n.a && !n.c && !n.g && !n.t && !n.n) return 'X';
    if (n.a >= n.c && n.a >= n.g && n.a >= n.t && n.a >= n.n) return 'A';
    if (n.c >= n.a && n.c >= n.g && n.c >= n.t && n.c >= n.n) return 'C';
    if (n.g >= n.a && n.g >= n.c && n.g >= n.t && n.g >= n.n) return 'G';
    if (n.t >= n.a && n.t >= n.c && n.t >= n.g && n.t >= n.n) return 'T';
    return 'N';
}
static int contain(u32 s1i, u32 s1, u32 e1i, u32 e1, u32 s2i, u32 s2, u32 e2i, u32 e2) {  // AG:1897-1902
    return s1i == s2i && e1i == e2i && s1 <= s2 && e1 >= e2;
}

static void walk(int coverage, int unit) {  // AG:1954-2204
    std::ofstream out((outdir + "/_pre_extended_contigs." + std::to_string(unit) + ".fa").c_str());
    for (Slot& sl : G) for (Node& n : sl.nodes) if (n.cid == NONE && n.cov < coverage) n.traversed = 1;  // AG:1904-1918
    u32 seq = 0, bsi = NONE, bso = NONE, bei = NONE, beo = NONE;
    for (u32 cp = 0; cp < G.size();) {
        for (u32 ip = 0; ip < G[cp].nodes.size(); ip++) {
            if (G[cp].nodes[ip].traversed) continue;
            Contig c; c.extended = 0; c.sid = 0; c.soff = cp; c.sid0 = G[cp].nodes[ip].mid; c.soff0 = G[cp].nodes[ip].moff;
            u32 p = cp, it = ip, pb = 0, ib = 0;  // current / "Bak" cursor
            int mode = 1;                            // kMerTag
            std::string tail;
            while ((mode == 1 && G[p].nodes[it].traversed == 0) || mode == 0) {
                if (mode == 0) { c.bases += G[p].cm[it].base; g_chain_steps++; }
                else { char b = consensus(G[p].nodes[it]); c.bases += (b != 'X') ? b : G[p].base; }
                if ((mode == 1 && G[p].nodes[it].coff != NONE) || mode == 0) c.extended = 1;
                if (mode == 1) {
                    Node& n = G[p].nodes[it];
                    n.traversed = 1; tail = n.s;
                    u32 cnt = 0, pick = NONE;
                    for (u32 e = 0; e < n.next.size(); e++)
                        if (n.next[e].id != NONE && G[n.next[e].off].nodes[n.next[e].item].traversed == 0) { pick = e; cnt++; }
                    if (cnt == 1) { pb = n.next[pick].off; ib = n.next[pick].item; p = pb; it = ib; mode = 1; }
                    else if (G[p].cm.size() == 1 && G[p].cm[0].nid != NONE) { pb = G[p].cm[0].noff; ib = G[p].cm[0].nitem; p = pb; it = ib; mode = 0; }
                    else mode = -1;
                } else {
                    const CMer& m = G[p].cm[it];
                    if (m.nid != NONE) { pb = m.noff; ib = m.nitem; p = pb; it = ib; mode = 0; }
                    else {
                        u32 live = 0, item = NONE, cnt = 0, pick = NONE;
                        for (u32 j = 0; j < G[p].nodes.size(); j++) if (G[p].nodes[j].traversed == 0) { live++; item = j; }
                        if (live == 1) {
                            const Node& n = G[p].nodes[item];
                            for (u32 e = 0; e < n.next.size(); e++)
                                if (n.next[e].id != NONE && G[n.next[e].off].nodes[n.next[e].item].traversed == 0) { cnt++; pick = e; }
                        }
                        if (cnt == 1) {
                            const Node& n = G[p].nodes[item];
                            pb = n.next[pick].off; ib = n.next[pick].item; p = pb; it = ib;
                            mode = G[p].nodes[it].traversed == 0 ? 1 : -2;
                        } else mode = -2;
                    }
                }
            }
            c.eid = 0; c.eoff = (mode == 1) ? pb : p;
            if (mode == 1 || mode == -1) { c.eid0 = G[p].nodes[it].mid; c.eoff0 = G[p].nodes[it].moff; }
            else { c.eid0 = NONE; c.eoff0 = NONE; }
            if (mode == -1 || mode == 1) {
                for (size_t j = 1; j < tail.size(); j++) c.bases += tail[j];
                c.eoff = (u32)(c.eoff + tail.size() - 1);      // size_t arithmetic truncated to u32 (AG:2170-2171)
                c.eoff0 = (u32)(c.eoff0 + tail.size() - 1);
            }
            g_walks++; g_walk_bases += (long)c.bases.size(); if ((long)c.bases.size() > g_walk_max) g_walk_max = (long)c.bases.size();
            if (!contain(bsi, bso, bei, beo, c.sid, c.soff, c.eid, c.eoff)) {
                g_emitted++; g_emitted_bases += (long)c.bases.size();
                out << ">" << seq++ << ", " << c.extended << ", " << c.sid << ", " << c.soff << ", " << c.eid << ", " << c.eoff << ", "
                    << c.sid0 << ", " << c.soff0 << ", " << c.eid0 << ", " << c.eoff0 << " \n";
                wrap60(out, c.bases);
                bsi = c.sid; bso = c.soff; bei = c.eid; beo = c.eoff;
                contigs.push_back(c);   // what extdContigs2 reloads from the file (AG:2258-2294)
            }
        }
        if (beo - bso > 100000) { if (0 == bei && cp + 1000 < beo) cp += 1000; else cp++; }  // AG:2194-2202
        else cp++;
    }
}

static void dedupJoin() {  // AG:2296-2380
    int n = (int)contigs.size();
    for (int a = 0; a < n; a++) {
        if (contigs[a].extended != 1) continue;
        for (int b = a + 1; b < n; b++) {
            if (contain(contigs[a].sid, contigs[a].soff, contigs[a].eid, contigs[a].eoff, contigs[b].sid, contigs[b].soff, contigs[b].eid, contigs[b].eoff)) contigs[b].extended = 2;
            else if (contigs[a].eid != contigs[b].sid || contigs[a].eoff < contigs[b].soff) break;
        }
    }
    for (int a = n - 1; a != -1; a--) {
        if (contigs[a].extended != 1) continue;
        for (int b = a - 1; b != -1; b--) {
            if (contain(contigs[a].sid, contigs[a].soff, contigs[a].eid, contigs[a].eoff, contigs[b].sid, contigs[b].soff, contigs[b].eid, contigs[b].eoff)) contigs[b].extended = 2;
            else if (contigs[b].eid != contigs[a].sid || contigs[b].eoff < contigs[a].soff) break;
        }
    }
    for (int a = 0; a < n; a++) {
        while (contigs[a].extended == 1) {
            int hits = 0, last = -1;
            for (int b = a + 1; b < n; b++) {
                if (contigs[b].extended == 2) continue;
                if (contigs[a].eoff >= contigs[b].soff) { hits++; last = b; }
                else break;
            }
            if (hits != 1) break;
            Contig snap = contigs[last];
            contigs[last].extended = 2;
            for (u32 i = contigs[a].eoff - snap.soff + 1; i < snap.bases.size(); i++) contigs[a].bases += snap.bases[i];
            contigs[a].eid = snap.eid; contigs[a].eoff = snap.eoff; contigs[a].eid0 = snap.eid0; contigs[a].eoff0 = snap.eoff0;
        }
    }
}

static int overlap(u32 x1, u32 y1, u32 x2, u32 y2) {  // AG:2388-2394
    return (x1 <= x2 && x2 <= y1 && y1 <= y2 && (int)y1 - (int)x2 > 0) || (x2 <= x1 && x1 <= y2 && y2 <= y1 && (int)y2 - (int)x1 > 0) ||
           (x1 <= x2 && x2 <= y2 && y2 <= y1 && (int)y2 - (int)x2 > 0) || (x2 <= x1 && x1 <= y1 && y1 <= y2 && (int)y1 - (int)x1 > 0);
}

static void scaffold(int unit) {  // AG:2396-2464
    std::vector<std::string> sc;
    for (u32 cp = 0; cp < contigs.size(); cp++) {
        if (!(contigs[cp].sid != NONE && contigs[cp].extended == 1)) continue;
        sc.push_back(contigs[cp].bases);
        contigs[cp].sid = NONE;
        int cont = 1;
        while (contigs[cp].sid0 == contigs[cp].eid0 && cont) {
            cont = 0;
            for (u32 c0 = cp + 1; c0 < contigs.size(); c0++) {
                const Contig& a = contigs[cp]; Contig& b = contigs[c0];
                if (!(c0 != cp && a.eid0 == b.sid && b.sid == b.eid && overlap(a.soff0, a.eoff0, b.soff, b.eoff) && b.extended == 1)) continue;
                if (b.soff > a.eoff) {
                    u32 gap = b.soff - a.eoff - 1;
                    int covered = 0;
                    for (u32 i = 0; i < gap; i++) if (!G[a.eoff + i + 1].nodes.empty() || !G[a.eoff + i + 1].cm.empty()) covered++;
                    if ((gap != 0 && (double)covered / gap >= 0.5) || gap == 0) { for (u32 i = 0; i < gap; i++) sc.back() += G[a.eoff + i + 1].base; }
                    else continue;
                }
                sc.back() += b.bases;
                b.sid = NONE;
                cp = c0; cont = 1;
                break;
            }
        }
    }
    std::ofstream out((outdir + "/_extended_contigs." + std::to_string(unit) + ".fa").c_str());
    for (size_t i = 0; i < sc.size(); i++) { out << ">" << i << "\n"; wrap60(out, sc[i]); }
}

// Optional node-table dump for node-level checks of the device build (text, one line per node).
static void dumpNodes(int unit) {
    std::ofstream out((outdir + "/_nodes." + std::to_string(unit) + ".txt").c_str());
    for (u32 p = 0; p < G.size(); p++)
        for (u32 i = 0; i < G[p].nodes.size(); i++) {
            const Node& n = G[p].nodes[i];
            std::vector<std::pair<u32, u32>> e;
            for (const Edge& x : n.next) e.push_back({x.off, x.item});
            std::sort(e.begin(), e.end());
            out << p << " " << i << " " << n.cov << " " << n.a << " " << n.c << " " << n.g << " " << n.t << " " << n.n << " " << n.cid << " " << n.coff << " "
                << n.cid0 << " " << n.coff0 << " " << n.mid << " " << n.moff << " [" << n.s << "]";
            for (auto& x : e) out << " " << x.first << ":" << x.second;
            out << "\n";
        }
}

int main(int argc, char** argv) {
    std::string dir = ".", contigFile, genomeFile;
    int k = 5, iv = 50, cov = 20, part = 1, first = 0, last = -1, dump = 0, prepare = 1, prepare_only = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--dir") dir = argv[++i];
        else if (a == "--out") outdir = argv[++i];
        else if (a == "--first") first = atoi(argv[++i]);
        else if (a == "--last") last = atoi(argv[++i]);
        else if (a == "--dump-nodes") dump = 1;
        else if (a == "--no-prepare") prepare = 0;
        else if (a == "--prepare-only") prepare_only = 1;
        else { fprintf(stderr, "ag_oracle: unknown option %s\n", a.c_str()); return 2; }
    }
    if (chdir(dir.c_str()) != 0) die("CANNOT OPEN FILE!");
    {   // parameters exactly as the reference reloads them on --resume (AG:4752-4753): tmp/_command.txt, one token per line
        std::ifstream cmd("tmp/_command.txt");
        if (!cmd.is_open()) die("CANNOT OPEN FILE!");
        std::string key, val;
        while (std::getline(cmd, key)) {
            if (key == "--misassemblyRemoval" || key == "--fastMap" || key == "--ratioCheck" || key == "--uniqueExtension" || key == "--iterativeMap") continue;
            if (!std::getline(cmd, val)) break;
            if (key == "--kMer") k = atoi(val.c_str());
            else if (key == "--insertVariation") iv = atoi(val.c_str());
            else if (key == "--coverage") cov = atoi(val.c_str());
            else if (key == "--part") part = atoi(val.c_str());
            else if (key == "--contig") contigFile = val;
            else if (key == "--genome") genomeFile = val;
        }
    }
    int units;
    if (prepare) { formalizeContigs(contigFile); units = formalizeGenome(genomeFile, part); }
    else { units = 0; while (access(("tmp/_genome." + std::to_string(units) + ".fa").c_str(), R_OK) == 0) units++; }
    if (prepare_only) return 0;
    if (last < 0 || last >= units) last = units - 1;
    for (int unit = first; unit <= last; unit++) {
        auto t0 = std::chrono::steady_clock::now();
        loadUnit(unit);
        {
            std::vector<Seq> chunks;
            loadChunks(chunks);
            loadContigAlignments(chunks, unit);
            threadContigs(chunks, unit);
        }
        auto t1 = std::chrono::steady_clock::now();
        loadReads(unit, k, iv);
        auto t2 = std::chrono::steady_clock::now();
        if (dump) dumpNodes(unit);
        walk(cov, unit);
        dedupJoin();
        scaffold(unit);
        auto t3 = std::chrono::steady_clock::now();
        size_t nn = 0, ne = 0; for (auto& s : G) { nn += s.nodes.size(); for (auto& n : s.nodes) ne += n.next.size(); }
        fprintf(stderr, "oracle unit %d: bp=%zu nodes=%zu edges=%zu events=%ld  contigs %.3fs reads %.3fs extend %.3fs\n", unit, G.size(), nn, ne, g_events,
                std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), std::chrono::duration<double>(t3 - t2).count());
        fprintf(stderr, "  walks=%ld walk_bases=%ld max_walk=%ld chain_steps=%ld emitted=%ld emitted_bases=%ld\n", g_walks, g_walk_bases, g_walk_max, g_chain_steps, g_emitted, g_emitted_bases);
        contigs.clear(); G.clear(); lastSource = NONE; g_events = 0; g_walks = g_walk_bases = g_walk_max = g_chain_steps = g_emitted = g_emitted_bases = 0;
    }
    return 0;
}

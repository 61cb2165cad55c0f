// agsynth — deterministic synthetic workload generator for the AlignGraph hot path.
//
// TEST / BENCH INFRASTRUCTURE (not product code).  Writes everything the per-chromosome loop of the
// reference (AlignGraph.cpp:4765-4783) consumes when it is entered through `--resume`
// (AlignGraph.cpp:4748-4760): the four user inputs (genome / contigs / reads_1 / reads_2 FASTA), and, under
// tmp/, `_reads.fa`, one truth-derived SAM (`_reads_genome.N.bowtie`) and PSL (`_contigs_genome.N.psl`) per
// unit N, `_command.txt` and `_checkpoint.txt`.  No aligner is needed: alignments are derived from the
// simulated reference->target edit script, which is what Bowtie2/BLAT would report on error-free data.
//
// Shapes follow SURVEY.md §8(d): reference iid ACGT, target = reference with SNPs (and, in "mix" mode, small
// indels), contigs = tiles of the target (every other one reverse-complemented), PE reads from the target with
// N(mean, sd) inserts, random mate order, SAM flags 99/147 or 83/163, integer QNAMEs in file order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <sys/stat.h>

struct Rng {  // splitmix64 / xorshift — deterministic across platforms
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct Params {
    std::string out = "synth";
    long genome_bp = 200000;
    int chroms = 1;
    double coverage = 50;
    int readlen = 100;
    double insert_mean = 500, insert_sd = 50;
    uint64_t seed = 20260925;
    double snp = 0.01;
    double indel = 0.0;        // per-bp rate of small indels (target vs reference)
    double read_err = 0.0;     // substitution errors in reads
    double n_rate = 0.0;       // 'N' bases in reads
    double softclip = 0.0;     // fraction of mates reported with soft clips
    double multi = 0.0;        // fraction of pairs with a second (-k) alignment record
    double unaligned = 0.0;    // fraction of pairs emitted as unaligned records
    double lowqual = 0.0;      // fraction of pairs whose CIGAR fails the 0.6 filter (heavy soft clip)
    long contig_len = 10000, contig_gap = 1000;
    int kmer = 5, cov = 20, part = 1, ivar = 50;
    int dist_low = 0, dist_high = 1500;
    int header = 0;            // write @SQ header lines into the SAM
    int misasm = 0;
    int user_reads = 1;        // 0: skip reads_1.fa / reads_2.fa (only a fresh run needs them) to save disk on the big configs
};

static const char ACGT[5] = "ACGT";
static inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
    return c;
}
static std::string revcomp(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (auto& c : r) c = comp(c);
    return r;
}

struct Out {
    FILE* f = nullptr;
    std::vector<char> buf;
    void open(const std::string& p) {
        f = fopen(p.c_str(), "wb");
        if (!f) { fprintf(stderr, "agsynth: cannot open %s\n", p.c_str()); exit(2); }
        buf.reserve(1 << 22);
    }
    void flush() { if (!buf.empty()) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); } }
    void put(const char* s, size_t n) { buf.insert(buf.end(), s, s + n); if (buf.size() > (1u << 22) - 4096) flush(); }
    void put(const std::string& s) { put(s.data(), s.size()); }
    void putc_(char c) { buf.push_back(c); }
    void num(long v) { char t[32]; int n = snprintf(t, sizeof t, "%ld", v); put(t, n); }
    void close() { if (f) { flush(); fclose(f); f = nullptr; } }
};

static void write_fasta(Out& o, const std::string& name, const std::string& seq) {
    o.putc_('>'); o.put(name); o.putc_('\n');
    for (size_t i = 0; i < seq.size(); i += 60) { o.put(seq.data() + i, std::min<size_t>(60, seq.size() - i)); o.putc_('\n'); }
}

// CIGAR + POS (0-based ref position of first M base) of target interval [a, a+len) through map t2r
// (t2r[i] = reference position of target base i, or -1 for an inserted base).
struct Aln { std::string cigar; long pos = -1; long ref_end = -1; int m = 0; bool ok = false; };
static Aln make_aln(const std::vector<int32_t>& t2r, long a, int len, int clip5, int clip3) {
    Aln r;
    long lo = a + clip5, hi = a + len - clip3;  // aligned window in target coords
    while (lo < hi && t2r[lo] < 0) lo++;
    while (hi > lo && t2r[hi - 1] < 0) hi--;
    if (hi - lo < 1) return r;
    std::vector<std::pair<char, int>> ops;
    auto push = [&](char c, int n) { if (n <= 0) return; if (!ops.empty() && ops.back().first == c) ops.back().second += n; else ops.push_back({c, n}); };
    push('S', (int)(lo - a));
    long prev = -1;
    for (long i = lo; i < hi; i++) {
        if (t2r[i] < 0) { push('I', 1); continue; }
        if (prev >= 0 && t2r[i] != prev + 1) push('D', (int)(t2r[i] - prev - 1));
        push('M', 1); r.m++;
        prev = t2r[i];
    }
    push('S', (int)(a + len - hi));
    r.pos = t2r[lo]; r.ref_end = prev;
    char t[24];
    for (auto& op : ops) { int n = snprintf(t, sizeof t, "%d%c", op.second, op.first); r.cigar.append(t, n); }
    r.ok = true;
    return r;
}

int main(int argc, char** argv) {
    Params p;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--out") p.out = val();
        else if (a == "--genome-bp") p.genome_bp = atol(val());
        else if (a == "--chroms") p.chroms = atoi(val());
        else if (a == "--coverage") p.coverage = atof(val());
        else if (a == "--readlen") p.readlen = atoi(val());
        else if (a == "--insert-mean") p.insert_mean = atof(val());
        else if (a == "--insert-sd") p.insert_sd = atof(val());
        else if (a == "--seed") p.seed = strtoull(val(), nullptr, 10);
        else if (a == "--snp") p.snp = atof(val());
        else if (a == "--indel") p.indel = atof(val());
        else if (a == "--read-err") p.read_err = atof(val());
        else if (a == "--n-rate") p.n_rate = atof(val());
        else if (a == "--softclip") p.softclip = atof(val());
        else if (a == "--multi") p.multi = atof(val());
        else if (a == "--unaligned") p.unaligned = atof(val());
        else if (a == "--lowqual") p.lowqual = atof(val());
        else if (a == "--contig-len") p.contig_len = atol(val());
        else if (a == "--contig-gap") p.contig_gap = atol(val());
        else if (a == "--kmer") p.kmer = atoi(val());
        else if (a == "--cov") p.cov = atoi(val());
        else if (a == "--part") p.part = atoi(val());
        else if (a == "--ivar") p.ivar = atoi(val());
        else if (a == "--header") p.header = atoi(val());
        else if (a == "--misasm") p.misasm = atoi(val());
        else if (a == "--user-reads") p.user_reads = atoi(val());
        else { fprintf(stderr, "agsynth: unknown option %s\n", a.c_str()); return 2; }
    }
    mkdir(p.out.c_str(), 0755);
    mkdir((p.out + "/tmp").c_str(), 0755);
    Rng rng(p.seed);

    const int L = p.readlen;
    long chr_bp = p.genome_bp / p.chroms;
    Out genome_fa, contigs_fa, r1, r2, rall;
    genome_fa.open(p.out + "/genome.fa");
    contigs_fa.open(p.out + "/contigs.fa");
    r1.open(p.out + "/reads_1.fa");
    r2.open(p.out + "/reads_2.fa");
    rall.open(p.out + "/tmp/_reads.fa");

    long pair_id = 0, contig_idx = 0, chunk_idx = 0, unit_base = 0;
    long n_units = 0;
    for (int c = 0; c < p.chroms; c++) {
        // ---- reference chromosome and diverged target -------------------------------------------------
        std::string ref(chr_bp, 'A');
        for (long i = 0; i < chr_bp; i++) ref[i] = ACGT[rng.below(4)];
        write_fasta(genome_fa, "chr" + std::to_string(c + 1), ref);
        std::string tgt; tgt.reserve(chr_bp + chr_bp / 50);
        std::vector<int32_t> t2r; t2r.reserve(chr_bp + chr_bp / 50);
        for (long i = 0; i < chr_bp; i++) {
            double u = rng.uni();
            if (p.indel > 0 && i > 50 && i < chr_bp - 50 && u < p.indel) {
                int n = 1 + (int)rng.below(3);
                if (rng.below(2)) { for (int j = 0; j < n; j++) { tgt.push_back(ACGT[rng.below(4)]); t2r.push_back(-1); } }
                else { i += n; }
                if (rng.below(8) == 0) {  // occasionally an insertion immediately followed by a deletion
                    tgt.push_back(ACGT[rng.below(4)]); t2r.push_back(-1); i += 1 + (long)rng.below(3);
                }
                if (i >= chr_bp) break;
            }
            char b = ref[i];
            if (rng.uni() < p.snp) b = ACGT[(std::strchr(ACGT, b) - ACGT + 1 + rng.below(3)) & 3];
            tgt.push_back(b); t2r.push_back((int32_t)i);
        }
        const long T = (long)tgt.size();

        // ---- unit boundaries (formalizeGenome, AlignGraph.cpp:3395-3409) ------------------------------
        std::vector<long> ustart;  // reference start of every unit of this chromosome
        {
            long q = 1; ustart.push_back(0);
            long chunk = chr_bp / p.part;
            for (long cp = 0; cp < chr_bp; cp++)
                if (cp != chr_bp - 1 && ((cp + 1) % chunk == 0 && q < p.part)) { ustart.push_back(cp + 1); q++; }
        }
        int nu = (int)ustart.size();
        auto unit_of = [&](long refpos) { int u = (int)(std::upper_bound(ustart.begin(), ustart.end(), refpos) - ustart.begin()) - 1; return u; };
        auto unit_end = [&](int u) { return u + 1 < nu ? ustart[u + 1] : chr_bp; };
        std::vector<Out> sam(nu), psl(nu);
        for (int u = 0; u < nu; u++) {
            sam[u].open(p.out + "/tmp/_reads_genome." + std::to_string(unit_base + u) + ".bowtie");
            psl[u].open(p.out + "/tmp/_contigs_genome." + std::to_string(unit_base + u) + ".psl");
            if (p.header) {
                sam[u].put("@HD\tVN:1.0\tSO:unsorted\n");
                sam[u].put("@SQ\tSN:" + std::to_string(unit_base + u) + "\tLN:" + std::to_string(unit_end(u) - ustart[u]) + "\n");
                sam[u].put("@PG\tID:bowtie2\tPN:bowtie2\tVN:synthetic\n");
            }
        }

        // ---- contigs: tiles of the target --------------------------------------------------------------
        for (long s = std::max(0L, p.contig_gap / 2); s + 300 < T; s += p.contig_len + p.contig_gap) {   // a negative gap makes the tiles overlap
            long e = std::min(T - 1, s + p.contig_len);
            // vary length a little so chunk sizes differ
            e = std::max(s + 250, e - (long)rng.below(p.contig_len / 10 + 1));
            long a = s, b = e;
            while (a < b && t2r[a] < 0) a++;
            while (b > a && t2r[b - 1] < 0) b--;
            if (b - a <= 200) continue;
            std::string seq = tgt.substr(a, b - a);
            bool rc = (contig_idx & 1);
            write_fasta(contigs_fa, "ctg" + std::to_string(contig_idx), rc ? revcomp(seq) : seq);
            // formalizeInput cuts contigs of >= 1,000,000 bp into chunks named "chunk.contig" (AlignGraph.cpp:3277-3293); BLAT then reports
            // one record per chunk, in chunk coordinates
            const long n = b - a;
            std::vector<std::pair<long, long>> chunks;   // [s, e) in FILE coordinates (the file holds the reverse complement when rc)
            if (n < 1000000) chunks.push_back({0, n});
            else { long cs = 0; for (long cpp = 0; cpp < n; cpp++) if ((cpp + 1) % 1000000 == 0 && cpp < n - 1 - 60) { chunks.push_back({cs, cpp + 1}); cs = cpp + 1; } chunks.push_back({cs, n}); }
            for (auto& chk : chunks) {
                long ta = rc ? a + (n - chk.second) : a + chk.first, tb = rc ? a + (n - chk.first) : a + chk.second;   // target interval of the chunk
                long fa = ta, fb = tb;
                while (fa < fb && t2r[fa] < 0) fa++;
                while (fb > fa && t2r[fb - 1] < 0) fb--;
                long this_chunk = chunk_idx++;
                if (fb - fa < 1) continue;
                int u = unit_of(t2r[fa]);
                if (unit_of(t2r[fb - 1]) != u) continue;
                // PSL blocks = maximal runs mapped to consecutive reference positions (query offsets relative to the chunk, on the
                // strand the chunk aligns with)
                std::vector<long> bs, qs, ts;
                long qins = 0, nqins = 0, tins = 0, ntins = 0, matches = 0;
                long i = fa;
                while (i < fb) {
                    if (t2r[i] < 0) { i++; continue; }
                    long j = i;
                    while (j + 1 < fb && t2r[j + 1] == t2r[j] + 1) j++;
                    bs.push_back(j - i + 1); qs.push_back(i - ta); ts.push_back(t2r[i] - ustart[u]);
                    matches += j - i + 1;
                    i = j + 1;
                }
                for (size_t k = 1; k < bs.size(); k++) {
                    long qg = qs[k] - (qs[k - 1] + bs[k - 1]), tg = ts[k] - (ts[k - 1] + bs[k - 1]);
                    if (qg > 0) { nqins++; qins += qg; }
                    if (tg > 0) { ntins++; tins += tg; }
                }
                Out& o = psl[u];
                long qsize = tb - ta;
                o.num(matches); o.put("\t0\t0\t0\t"); o.num(nqins); o.putc_('\t'); o.num(qins); o.putc_('\t');
                o.num(ntins); o.putc_('\t'); o.num(tins); o.putc_('\t'); o.putc_(rc ? '-' : '+'); o.putc_('\t');
                o.num(this_chunk); o.putc_('.'); o.num(contig_idx); o.putc_('\t'); o.num(qsize); o.putc_('\t'); o.num(qs.front()); o.putc_('\t'); o.num(qs.back() + bs.back());
                o.put("\t0\t"); o.num(unit_end(u) - ustart[u]); o.putc_('\t'); o.num(ts.front()); o.putc_('\t');
                o.num(ts.back() + bs.back()); o.putc_('\t'); o.num((long)bs.size()); o.putc_('\t');
                for (long v : bs) { o.num(v); o.putc_(','); } o.putc_('\t');
                for (long v : qs) { o.num(v); o.putc_(','); } o.putc_('\t');
                for (long v : ts) { o.num(v); o.putc_(','); } o.putc_('\n');
            }
            contig_idx++;
        }

        // ---- paired-end reads ----------------------------------------------------------------------------
        long n_pairs = (long)(chr_bp * p.coverage / (2.0 * L));
        std::string m1, m2;
        for (long n = 0; n < n_pairs; n++, pair_id++) {
            long ins = (long)std::llround(p.insert_mean + p.insert_sd * rng.normal());
            if (ins < L + 10) ins = L + 10;
            if (ins >= T) ins = T - 1;
            long f = (long)rng.below((uint64_t)(T - ins));
            std::string fwd = tgt.substr(f, L);
            std::string rev = revcomp(tgt.substr(f + ins - L, L));
            for (std::string* s : {&fwd, &rev})
                for (auto& ch : *s) {
                    if (p.read_err > 0 && rng.uni() < p.read_err) ch = ACGT[rng.below(4)];
                    if (p.n_rate > 0 && rng.uni() < p.n_rate) ch = 'N';
                }
            bool fwd_is_1 = rng.below(2) == 0;
            const std::string& s1 = fwd_is_1 ? fwd : rev;
            const std::string& s2 = fwd_is_1 ? rev : fwd;
            for (Out* o : {&r1, &rall}) { if (o != &rall && !p.user_reads) continue; o->putc_('>'); o->num(pair_id); o->putc_('\n'); o->put(s1); o->putc_('\n'); }
            for (Out* o : {&r2, &rall}) { if (o != &rall && !p.user_reads) continue; o->putc_('>'); o->num(pair_id); o->putc_('\n'); o->put(s2); o->putc_('\n'); }

            int c5f = 0, c3f = 0, c5r = 0, c3r = 0;
            if (p.softclip > 0 && rng.uni() < p.softclip) { c5f = (int)rng.below(12); c3f = (int)rng.below(12); }
            if (p.softclip > 0 && rng.uni() < p.softclip) { c5r = (int)rng.below(12); c3r = (int)rng.below(12); }
            if (p.lowqual > 0 && rng.uni() < p.lowqual) { c5f = L / 4 + (int)rng.below(L / 4); c3f = L / 5; }
            Aln af = make_aln(t2r, f, L, c5f, c3f);
            Aln ar = make_aln(t2r, f + ins - L, L, c5r, c3r);  // reverse mate, reported on the forward strand
            bool unal = p.unaligned > 0 && rng.uni() < p.unaligned;
            if (!af.ok || !ar.ok) continue;
            int u = unit_of(af.pos);
            if (unit_of(af.ref_end) != u || unit_of(ar.pos) != u || unit_of(ar.ref_end) != u) continue;
            long ub = ustart[u];
            Out& o = sam[u];
            auto rec = [&](int flag, const Aln* a, long mpos, long tlen) {
                o.num(pair_id); o.putc_('\t'); o.num(flag); o.putc_('\t');
                if (!a) { o.put("*\t0\t0\t*\t*\t0\t0\t*\t*\tYT:Z:UP\n"); return; }
                o.num(unit_base + u); o.putc_('\t'); o.num(a->pos - ub + 1); o.put("\t44\t"); o.put(a->cigar);
                o.put("\t=\t"); o.num(mpos - ub + 1); o.putc_('\t'); o.num(tlen); o.put("\t*\t*\tAS:i:0\tYT:Z:CP\n");
            };
            if (unal) { rec(77, nullptr, 0, 0); rec(141, nullptr, 0, 0); continue; }
            auto emit_pair = [&](const Aln& F, const Aln& R) {
                if (fwd_is_1) { rec(99, &F, R.pos, ins); rec(147, &R, F.pos, -ins); }
                else          { rec(83, &R, F.pos, -ins); rec(163, &F, R.pos, ins); }
            };
            emit_pair(af, ar);
            if (p.multi > 0 && rng.uni() < p.multi) {
                // a second "-k" hit: half of them within one read length (exercises the duplicate rule,
                // AlignGraph.cpp:1650-1655), the rest elsewhere in the same unit
                long shift = rng.below(2) ? (long)rng.below(L) - L / 2 : (long)rng.below(5000) + L;
                long f2 = f + shift;
                if (f2 >= 0 && f2 + ins < T) {
                    Aln bf = make_aln(t2r, f2, L, 0, 0), br = make_aln(t2r, f2 + ins - L, L, 0, 0);
                    if (bf.ok && br.ok && unit_of(bf.pos) == u && unit_of(bf.ref_end) == u && unit_of(br.pos) == u && unit_of(br.ref_end) == u)
                        emit_pair(bf, br);
                }
            }
        }
        for (int u = 0; u < nu; u++) { sam[u].close(); psl[u].close(); }
        unit_base += nu; n_units += nu;
    }
    genome_fa.close(); contigs_fa.close(); r1.close(); r2.close(); rall.close();

    Out cmd; cmd.open(p.out + "/tmp/_command.txt");
    auto arg = [&](const std::string& k, const std::string& v) { cmd.put(k); cmd.putc_('\n'); cmd.put(v); cmd.putc_('\n'); };
    arg("--read1", "reads_1.fa"); arg("--read2", "reads_2.fa"); arg("--contig", "contigs.fa"); arg("--genome", "genome.fa");
    arg("--distanceLow", std::to_string(p.dist_low)); arg("--distanceHigh", std::to_string(p.dist_high));
    arg("--extendedContig", "extendedContigs.fa"); arg("--remainingContig", "remainingContigs.fa");
    arg("--kMer", std::to_string(p.kmer)); arg("--coverage", std::to_string(p.cov));
    arg("--insertVariation", std::to_string(p.ivar)); arg("--part", std::to_string(p.part));
    if (p.misasm) { cmd.put("--misassemblyRemoval\n"); }
    cmd.close();
    Out cp; cp.open(p.out + "/tmp/_checkpoint.txt"); cp.put("0\n"); cp.close();
    Out meta; meta.open(p.out + "/synth_meta.txt");
    meta.put("pairs " + std::to_string(pair_id) + "\nunits " + std::to_string(n_units) + "\ncontigs " + std::to_string(contig_idx) +
             "\ngenome_bp " + std::to_string(chr_bp * p.chroms) + "\nreadlen " + std::to_string(L) + "\nkmer " + std::to_string(p.kmer) + "\n");
    meta.close();
    return 0;
}

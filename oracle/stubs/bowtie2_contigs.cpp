// Stub `bowtie2` for read-vs-contig alignment — TEST INFRASTRUCTURE ONLY.
//
// removeMisassembly (AlignGraph.cpp:3821-3832) shells out to
//     bowtie2 -f --no-mixed -k 1 -p 8 -I lo -X hi --no-discordant -x tmp/_<id>_contigs -1 tmp/_reads_1.fa -2 tmp/_reads_2.fa --reorder
// Bowtie2 is not installed in this image, so the harness puts this stand-in on $PATH for BOTH implementations (the shell stub `bowtie2`
// dispatches here when the index name ends in "_contigs").  It places every mate by an exact 16-mer seed at the read's start and an ungapped
// comparison over the whole read (>= 95 % identity, either strand, first hit in (contig, offset) order) and reports a pair only when both
// mates land on the same contig, one per strand — enough to give removeMisassembly a realistic per-base coverage pile-up, and identical for
// the reference and for this repo.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

static void load(const std::string& path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    std::ifstream in(path.c_str());
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '>') { names.push_back(line.substr(1)); seqs.emplace_back(); }
        else if (!seqs.empty()) seqs.back() += line;
    }
}
static std::string rc(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (auto& c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    return r;
}
struct Hit { int contig = -1; long pos = 0; int strand = 0; };

int main(int argc, char** argv) {
    std::string index, f1, f2;
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "-x")) index = argv[i + 1];
        else if (!strcmp(argv[i], "-1")) f1 = argv[i + 1];
        else if (!strcmp(argv[i], "-2")) f2 = argv[i + 1];
    }
    std::vector<std::string> cn, cs, n1, r1, n2, r2;
    load(index + ".fa", cn, cs);
    load(f1, n1, r1);
    load(f2, n2, r2);
    const size_t K = 16;
    std::unordered_map<std::string, std::vector<std::pair<int, long>>> idx;
    for (size_t c = 0; c < cs.size(); c++)
        for (size_t i = 0; i + K <= cs[c].size(); i++) idx[cs[c].substr(i, K)].push_back({(int)c, (long)i});
    auto place = [&](const std::string& read) -> Hit {
        Hit best;
        for (int strand = 0; strand < 2; strand++) {
            const std::string s = strand ? rc(read) : read;
            if (s.size() < K) continue;
            auto it = idx.find(s.substr(0, K));
            if (it == idx.end()) continue;
            for (auto& h : it->second) {
                if (h.second + (long)s.size() > (long)cs[h.first].size()) continue;
                long match = 0;
                for (size_t i = 0; i < s.size(); i++) match += s[i] == cs[h.first][h.second + i];
                if (match * 100 < (long)s.size() * 95) continue;
                if (best.contig < 0 || std::make_pair(h.first, h.second) < std::make_pair(best.contig, best.pos)) { best.contig = h.first; best.pos = h.second; best.strand = strand; }
            }
        }
        return best;
    };
    for (size_t c = 0; c < cs.size(); c++) printf("@SQ\tSN:%s\tLN:%zu\n", cn[c].c_str(), cs[c].size());
    const size_t n = std::min(r1.size(), r2.size());
    for (size_t p = 0; p < n; p++) {
        const Hit a = place(r1[p]), b = place(r2[p]);
        const char* q = n1[p].c_str();
        if (a.contig < 0 || b.contig != a.contig || a.strand == b.strand) {
            printf("%s\t77\t*\t0\t0\t*\t*\t0\t0\t*\t*\tYT:Z:UP\n%s\t141\t*\t0\t0\t*\t*\t0\t0\t*\t*\tYT:Z:UP\n", q, q);
            continue;
        }
        const int fa = a.strand ? 83 : 99, fb = a.strand ? 163 : 147;
        printf("%s\t%d\t%s\t%ld\t42\t%zuM\t=\t%ld\t0\t*\t*\tYT:Z:CP\n", q, fa, cn[a.contig].c_str(), a.pos + 1, r1[p].size(), b.pos + 1);
        printf("%s\t%d\t%s\t%ld\t42\t%zuM\t=\t%ld\t0\t*\t*\tYT:Z:CP\n", q, fb, cn[b.contig].c_str(), b.pos + 1, r2[p].size(), a.pos + 1);
    }
    return 0;
}

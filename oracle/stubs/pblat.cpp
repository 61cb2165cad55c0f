// Stub `pblat` / `blat` for the oracle harness — TEST INFRASTRUCTURE ONLY.
//
// The reference's refinement() (AlignGraph.cpp:2976) shells out to `pblat <db.fa> <query.fa> -noHead <out.psl>
// -fastMap -threads=8`.  BLAT is not installed in this image, so the harness puts this stand-in on $PATH for BOTH
// implementations: it reports, for every query, every ungapped placement in a database sequence (either strand)
// that is seeded by an exact 24-mer and has >= 90 % identity over the full query, as a single-block PSL line.
// That is enough for the containment test at AlignGraph.cpp:3059 and keeps both sides on identical aligner output.
// A query without any such placement gets LOCAL ungapped alignments instead (seed + X-drop extension), which is what removeMisassembly
// (AlignGraph.cpp:4003-4145) needs to see for chimeric contigs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <set>
#include <unordered_map>
#include <algorithm>

static void load(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    std::ifstream in(path);
    if (!in.is_open()) { fprintf(stderr, "stub pblat: cannot open %s\n", path); exit(1); }
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

int main(int argc, char** argv) {
    std::vector<const char*> pos;
    for (int i = 1; i < argc; i++) if (argv[i][0] != '-') pos.push_back(argv[i]);
    if (pos.size() < 3) { fprintf(stderr, "usage: pblat db.fa query.fa [-noHead] out.psl\n"); return 1; }
    std::vector<std::string> tn, ts, qn, qs;
    load(pos[0], tn, ts);
    load(pos[1], qn, qs);
    FILE* out = fopen(pos[2], "w");
    if (!out) return 1;
    const size_t K = 24;
    // index database 24-mers sampled at every position (hash -> list of (target, offset))
    std::unordered_multimap<std::string, std::pair<int, long>> idx;
    for (size_t t = 0; t < ts.size(); t++)
        for (size_t i = 0; i + K <= ts[t].size(); i += 1) idx.emplace(ts[t].substr(i, K), std::make_pair((int)t, (long)i));
    for (size_t q = 0; q < qs.size(); q++) {
        for (int strand = 0; strand < 2; strand++) {
            std::string s = strand ? rc(qs[q]) : qs[q];
            if (s.size() < K) continue;
            std::set<std::pair<int, long>> seen;
            bool placed = false;
            for (size_t off = 0; off + K <= s.size(); off += std::max<size_t>(K, s.size() / 16)) {
                auto range = idx.equal_range(s.substr(off, K));
                for (auto it = range.first; it != range.second; ++it) {
                    int t = it->second.first; long start = it->second.second - (long)off;
                    if (start < 0 || start + (long)s.size() > (long)ts[t].size()) continue;
                    if (!seen.insert({t, start}).second) continue;
                    long match = 0;
                    for (size_t i = 0; i < s.size(); i++) match += (s[i] == ts[t][start + i]);
                    if (match * 10 < (long)s.size() * 9) continue;
                    placed = true;
                    fprintf(out, "%ld\t%ld\t0\t0\t0\t0\t0\t0\t%c\t%s\t%zu\t0\t%zu\t%s\t%zu\t%ld\t%ld\t1\t%zu,\t0,\t%ld,\n", match,
                            (long)s.size() - match, strand ? '-' : '+', qn[q].c_str(), s.size(), s.size(), tn[t].c_str(),
                            ts[t].size(), start, start + (long)s.size(), s.size(), start);
                }
            }
            if (placed) continue;
            // no full-length placement: LOCAL ungapped alignments (what BLAT reports for a chimeric / partly foreign query).  Seeds every
            // K bases; each new diagonal is extended in both directions with an X-drop (match +1, mismatch -3, drop 30); segments of at
            // least 100 bases are reported as single-block PSL lines with their query interval (strand '-': interval in the reverse complement,
            // as BLAT does).  Deterministic: seeds left to right, hits in index order, one report per (target, diagonal, segment).
            std::set<std::pair<std::pair<int, long>, long>> done;   // ((target, diagonal), query start)
            for (size_t off = 0; off + K <= s.size(); off += K) {
                auto range = idx.equal_range(s.substr(off, K));
                std::vector<std::pair<int, long>> hits;
                for (auto it = range.first; it != range.second; ++it) hits.push_back(it->second);
                std::sort(hits.begin(), hits.end());
                for (auto& h : hits) {
                    const int t = h.first; const long diag = h.second - (long)off;
                    long qs = (long)off, qe = (long)off + (long)K;   // [qs, qe) in s; target = q + diag
                    { long score = 0, best = 0, i = qe; long be = qe;
                      while (i < (long)s.size() && i + diag < (long)ts[t].size()) { score += s[i] == ts[t][i + diag] ? 1 : -3; i++; if (score > best) { best = score; be = i; } if (score < best - 30) break; }
                      qe = be; }
                    { long score = 0, best = 0, i = qs - 1; long bs = qs;
                      while (i >= 0 && i + diag >= 0) { score += s[i] == ts[t][i + diag] ? 1 : -3; if (score > best) { best = score; bs = i; } if (score < best - 30) break; i--; }
                      qs = bs; }
                    if (qe - qs < 100) continue;
                    if (!done.insert({{t, diag}, qs}).second) continue;
                    long match = 0;
                    for (long i = qs; i < qe; i++) match += s[i] == ts[t][i + diag];
                    fprintf(out, "%ld\t%ld\t0\t0\t0\t0\t0\t0\t%c\t%s\t%zu\t%ld\t%ld\t%s\t%zu\t%ld\t%ld\t1\t%ld,\t%ld,\t%ld,\n", match, (qe - qs) - match, strand ? '-' : '+',
                            qn[q].c_str(), s.size(), qs, qe, tn[t].c_str(), ts[t].size(), qs + diag, qe + diag, qe - qs, qs, qs + diag);
                }
            }
        }
    }
    fclose(out);
    return 0;
}

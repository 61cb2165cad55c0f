// AlignGraph — drop-in command line of the B200 build.  Same flags, same tmp/ file contract, same progress lines and final
// FASTA as the reference's main() (AG:4696-4796); the per-chromosome loop body (AG:4768-4776) runs on the GPU through the C ABI
// of libaligngraph_b200 (include/aligngraph_b200.h).  Bowtie2 / BLAT are shelled out to with the reference's exact command lines
// (AG:3599-3610, AG:3648-3652, AG:2976-2980).
//
// Multi-GPU: units (chromosomes or --part slices) are independent (AG:4779-4781 clears all state); set AG_DEVICES=0,1,... to farm
// them round-robin over several GPUs (one context and one host thread per GPU).  Outputs are identical for any device count.
#include "../../include/aligngraph_b200.h"
#include "ag_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>
#include <mutex>
#include <atomic>

using std::cout;
using std::endl;
using std::string;

static const int kMax = 99999;  // MAX, AG:27

struct Options {
    string read1, read2, contig, genome, ext, rmn;
    int tagRead1 = 0, tagRead2 = 0, tagContig = 0, tagGenome = 0, tagExt = 0, tagRmn = 0, tagKMer = 0, tagLow = 0, tagHigh = 0, tagIv = 0, tagCov = 0,
        tagPart = 0, tagFastMap = 0, tagRatio = 0, tagUnique = 0, tagIter = 0, tagMis = 0, tagResume = 0;
    int k = 5, low = 0, high = kMax, iv = 50, cov = 20, part = 1;
};

static void usage() {  // AG:4304-4327, verbatim
    cout << "AlignGraph --read1 reads_1.fa --read2 reads_2.fa --contig contigs.fa --genome genome.fa --distanceLow distanceLow --distanceHigh distancehigh --extendedContig extendedContigs.fa --remainingContig remainingContigs.fa [--kMer k --insertVariation insertVariation --covereage coverage --part p --ratioCheck --iterativeMap --misassemblyRemoval --resume]" << endl;
    cout << "Inputs:" << endl;
    cout << "--read1 is the the first pair of PE DNA reads in fasta format" << endl;
    cout << "--read2 is the the second pair of PE DNA reads in fasta format" << endl;
    cout << "--contig is the initial contigs in fasta format" << endl;
    cout << "--genome is the reference genome in fasta format" << endl;
    cout << "--distanceLow is the lower bound of alignment distance between the first and second pairs of PE DNA reads (recommended: max{insert length - 1000, single read length})" << endl;
    cout << "--distanceHigh is the upper bound of alignment distance between the first and second pairs of PE DNA reads (recommended: insert length + 1000)" << endl;
    cout << "Outputs:" << endl;
    cout << "--extendedContig is the extended contig file in fasta format" << endl;
    cout << "--remainingContig is the not extended initial contig file in fasta format" << endl;
    cout << "Options:" << endl;
    cout << "--kMer is the k-mer size (default: 5)" << endl;
    cout << "--insertVariation is the small variation of insert length (default: 50)" << endl;
    cout << "--coverage is the minimum coverage to keep a path in de Bruijn graph (default: 20)" << endl;
    cout << "--part is the number of parts a chromosome is divided into when it is loaded to reduce memory requirement (default: 1)" << endl;
    cout << "--fastMap calls NUCMER to make fast but less sensitive and accurate contig alignment instead of BLAT (default: none)" << endl;
    cout << "--ratioCheck checks read alignment ratio to the reference beforehand and warns if the ratio is too low; may take a little more time (default: none)" << endl;
    cout << "--iterativeMap aligns reads to one chromosome and then another rather than directly to the genome, which increases sensitivity while loses precision (default: none)" << endl;
    cout << "--misassemblyRemoval detects and then breaks at or removes misassembed regions (default: none)" << endl;
    cout << "--resume resumes the previous unfinished running from several checkpoints (default: none)" << endl;
}

[[noreturn]] static void die_usage() { usage(); exit(-1); }
[[noreturn]] static void die(const string& msg) { cout << msg << endl; exit(-1); }

static bool readable(const string& p) { std::ifstream f(p.c_str()); return f.is_open(); }

// getParameters (AG:4329-4646): one token per line; a flag's value is the next line; integers must round-trip through atoi.
static void parse_command_file(const string& path, Options& o) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) die("CANNOT OPEN FILE!");
    std::vector<string> tok;
    string buf;
    while (in.good()) { std::getline(in, buf); if (buf[0] == 0) break; tok.push_back(buf); }
    int count = (int)tok.size(), i = -1;
    size_t pos = 0;
    auto value = [&]() -> string { return pos < tok.size() ? tok[pos++] : string(); };
    auto file_in = [&](int& tag, string& dst) {
        if (tag == 1 || i == count - 1) die_usage();
        dst = value();
        if (!readable(dst)) { cout << "CANNOT OPEN FILE!" << endl; die_usage(); }
        tag = 1;
    };
    auto file_out = [&](int& tag, string& dst) {
        if (tag == 1 || i == count - 1) die_usage();
        dst = value();
        std::ofstream f(dst.c_str());  // the reference opens (and truncates) the output here
        if (!f.is_open()) { cout << "CANNOT OPEN FILE!" << endl; die_usage(); }
        tag = 1;
    };
    auto integer = [&](int& tag, int& dst) {
        if (tag == 1 || i == count - 1) die_usage();
        string v = value();
        dst = atoi(v.c_str());
        std::stringstream ss; ss << dst;
        if (ss.str() != v) die_usage();
        tag = 1;
    };
    auto flag = [&](int& tag) { if (tag == 1) die_usage(); tag = 1; };
    while (pos < tok.size()) {
        buf = tok[pos++];
        i++;
        if (buf == "--read1") file_in(o.tagRead1, o.read1);
        else if (buf == "--read2") file_in(o.tagRead2, o.read2);
        else if (buf == "--contig") file_in(o.tagContig, o.contig);
        else if (buf == "--genome") file_in(o.tagGenome, o.genome);
        else if (buf == "--distanceLow") integer(o.tagLow, o.low);
        else if (buf == "--distanceHigh") integer(o.tagHigh, o.high);
        else if (buf == "--extendedContig") file_out(o.tagExt, o.ext);
        else if (buf == "--remainingContig") file_out(o.tagRmn, o.rmn);
        else if (buf == "--kMer") integer(o.tagKMer, o.k);
        else if (buf == "--insertVariation") integer(o.tagIv, o.iv);
        else if (buf == "--coverage") integer(o.tagCov, o.cov);
        else if (buf == "--part") integer(o.tagPart, o.part);
        else if (buf == "--fastMap") flag(o.tagFastMap);
        else if (buf == "--ratioCheck") flag(o.tagRatio);
        else if (buf == "--uniqueExtension") flag(o.tagUnique);
        else if (buf == "--iterativeMap") flag(o.tagIter);
        else if (buf == "--misassemblyRemoval") flag(o.tagMis);
        else if (buf == "--resume") { if (o.tagResume == 1 || count != 1) die_usage(); o.tagResume = 1; }
        else die_usage();
    }
}

// ---- external aligners, command lines verbatim (AG:3581-3656) ---------------------------------------------------------------
static void run_bowtie(int low, int high, int units, int iterative) {
    std::stringstream lo, hi; lo << low; hi << high;
    if (iterative == 1) {
        for (int u = 0; u < units; u++) {
            string n = std::to_string(u);
            string c = "bowtie2-build -f tmp/_genome." + n + ".fa tmp/_genome." + n + " > bowtie_doc.txt 2> bowtie_doc.txt";
            if (system(c.c_str())) {}
            c = "bowtie2 -f --no-mixed -k 5 -p 8 --local --mp 3,1 --rdg 2,1 --rfg 2,1 --score-min G,5,2 -I " + lo.str() + " -X " + hi.str() +
                " --no-discordant -x tmp/_genome." + n + " -1 tmp/_reads_1.fa -2 tmp/_reads_2.fa --reorder > tmp/_reads_genome." + n + ".bowtie 2> bowtie_doc.txt";
            if (system(c.c_str())) {}
        }
    } else {
        string c = "bowtie2-build -f tmp/_genome.fa tmp/_genome > bowtie_doc.txt 2> bowtie_doc.txt";
        if (system(c.c_str())) {}
        c = "bowtie2 -f --no-mixed -k 5 -p 8 --local --mp 3,1 --rdg 2,1 --rfg 2,1 --score-min G,5,2 -I " + lo.str() + " -X " + hi.str() +
            " --no-discordant -x tmp/_genome -1 tmp/_reads_1.fa -2 tmp/_reads_2.fa --reorder > tmp/_reads_genome.bowtie 2> bowtie_doc.txt";
        if (system(c.c_str())) {}
        ag_distribute_alignments("tmp", units);
    }
}
static bool blat_call(const string& db, const string& query, const string& out) {
    string c = "pblat " + db + " " + query + " -noHead " + out + " -fastMap -threads=8 > blat_doc.txt 2> blat_doc.txt";
    if (system(c.c_str()) != 0) {
        c = "blat " + db + " " + query + " -noHead " + out + " -fastMap > blat_doc.txt 2> blat_doc.txt";
        if (system(c.c_str()) != 0) return false;
    }
    return true;
}
static void run_blat(int units) {
    for (int u = 0; u < units; u++) {
        string n = std::to_string(u);
        if (!blat_call("tmp/_genome." + n + ".fa", "tmp/_contigs.fa", "tmp/_contigs_genome." + n + ".psl")) throw AgHostError{"BLAT CALL FAILED!"};
    }
}
// refinement's aligner call (AG:2957-2983).  Without pblat / blat on $PATH (or with AG_BUILTIN_CONTAINMENT=1) the library's own containment
// search produces the PSL: the reference would stop with BLAT CALL FAILED!, this build can finish the run on a host that has no BLAT.
static bool refine_blat(int unit, void* user) {
    string n = std::to_string(unit);
    const string db = "tmp/_extended_contigs." + n + ".fa", q = "tmp/_short_initial_contigs." + n + ".fa", out = "tmp/_short_initial_contigs_extended_contigs." + n + ".psl";
    if (!getenv("AG_BUILTIN_CONTAINMENT") && blat_call(db, q, out)) return true;
    ag_ctx* ctx = (ag_ctx*)user;
    return ctx && ag_containment_search_files(ctx, db.c_str(), q.c_str(), out.c_str()) == 0;
}

int main(int argc, char* argv[]) {
    cout << "AlignGraph: algorithm for secondary de novo genome assembly guided by closely related references" << endl;
    cout << "By Ergude Bao, CS Department, UC-Riverside. All Rights Reserved" << endl << endl;
    time_t start = time(NULL), startAlign, endAlign;
    ag_tune_malloc();   // the CLI owns its process: keep the large staging blocks on the heap from unit to unit
    {
        std::ofstream w("command.txt");
        if (!w.is_open()) { cout << "CANNOT OPEN FILE!" << endl; return 0; }
        for (int i = 1; i < argc; i++) w << argv[i] << endl;
    }
    Options o;
    parse_command_file("command.txt", o);
    std::vector<string> contig_ids, genome_ids;
    int units = 0, cp = 0;
    std::ofstream wcp;
    try {
        if (o.tagResume == 0) {
            if (o.tagRead1 == 0 || o.tagRead2 == 0 || o.tagContig == 0 || o.tagGenome == 0 || o.tagExt == 0 || o.tagRmn == 0 || o.k <= 0 || o.tagLow == 0 ||
                o.tagHigh == 0 || o.low > o.high || o.low < 0 || o.iv < 0 || o.part < 1 || o.part > 10 || o.k > ag_max_read_length(o.read1) ||
                o.k > ag_max_read_length(o.read2)) { usage(); return 0; }
            if (o.tagFastMap) die("--fastMap (NUCMER) is not supported by the B200 build; use the default BLAT mapping");
            if (system("bowtie2 -h > bowtie_doc.txt 2> bowtie_doc.txt") != 0) die("BOWTIE2 CALL FAILED!");
            if (system("test -d \"tmp\"; t=$?; if [ $t -eq 1 ]; then mkdir tmp; fi")) {}
            { std::ofstream w("tmp/_command.txt"); for (int i = 1; i < argc; i++) w << argv[i] << endl; }
            wcp.open("tmp/_checkpoint.txt");
            ag_formalize_reads(o.read1, o.read2, "tmp");
            ag_formalize_contigs(o.contig, "tmp", contig_ids);
            units = ag_formalize_genome(o.genome, "tmp", o.part, genome_ids);
            startAlign = time(NULL);
            {   // parallelMap (AG:3720-3735): the read and the contig alignment jobs run side by side
                // a failure inside either thread is reported from the main thread after both have joined (message + exit(-1) as in the reference)
                string err0, err1;
                std::thread t0([&] { try { run_bowtie(o.low, o.high, units, o.tagIter); } catch (const AgHostError& e) { err0 = e.msg; } catch (...) { err0 = "BOWTIE2 CALL FAILED!"; } });
                std::thread t1([&] { try { run_blat(units); } catch (const AgHostError& e) { err1 = e.msg; } catch (...) { err1 = "BLAT CALL FAILED!"; } });
                t0.join(); t1.join();
                if (!err0.empty()) die(err0);
                if (!err1.empty()) die(err1);
            }
            endAlign = time(NULL);
            cout << "(0) Alignment finished" << endl;
            wcp << "0" << endl;
        } else {
            {   // getCheckpoint (AG:4653-4680): the last line wins
                std::ifstream r("tmp/_checkpoint.txt");
                if (!r.is_open()) die("CANNOT OPEN FILE!");
                cp = -1; string s;
                while (r.good()) { std::getline(r, s); if (s[0] == 0) break; cp = atoi(s.c_str()); }
                if (cp == -1) die("NOT REACHED CHECKPOINT. PLEASE RERUN!");
            }
            o = Options(); o.tagResume = 1;
            parse_command_file("tmp/_command.txt", o);
            if (o.tagFastMap) die("--fastMap (NUCMER) is not supported by the B200 build; use the default BLAT mapping");
            cout << "RESUMED SUCCESSFULLY :-)" << endl;
            wcp.open("tmp/_checkpoint.txt", std::ios::app);
            ag_formalize_contigs(o.contig, "tmp", contig_ids);
            units = ag_formalize_genome(o.genome, "tmp", o.part, genome_ids);
            startAlign = endAlign = time(NULL);
        }
        if (o.tagRatio == 1) {
            double ratio = ag_check_ratio("tmp", units);
            cout << " - " << ratio * 100 << "% reads aligned ";
            if (ratio < 0.25) cout << "(warning: ratio below 25%; hard to guarantee good results)" << endl; else cout << endl;
        }
    } catch (const AgHostError& e) { die(e.msg); }

    // ---- the hot loop (AG:4765-4783) on the GPU(s) ---------------------------------------------------------------------------
    std::vector<int> devices;
    if (const char* env = getenv("AG_DEVICES")) { std::stringstream ss(env); string t; while (std::getline(ss, t, ',')) if (!t.empty()) devices.push_back(atoi(t.c_str())); }
    if (devices.empty()) devices.push_back(0);
    if (cp < units) {
        // more units than GPUs: two contexts per GPU, so that one unit's text staging and host post passes run under the other's kernels
        // (AG_CONTEXTS_PER_DEVICE overrides; outputs do not depend on it)
        int per_dev = (units - cp) > (int)devices.size() ? 2 : 1;
        if (const char* e = getenv("AG_CONTEXTS_PER_DEVICE")) { const int v = atoi(e); if (v >= 1 && v <= 4) per_dev = v; }
        { const std::vector<int> base = devices; for (int r = 1; r < per_dev; r++) devices.insert(devices.end(), base.begin(), base.end()); }
        std::vector<ag_ctx*> ctxs;
        for (int d : devices) {
            ag_params p; p.k = o.k; p.insert_variation = o.iv; p.coverage = o.cov; p.device = d;
            ag_ctx* c = nullptr;
            if (ag_create(&p, &c) != 0) die(ag_create_error());
            ctxs.push_back(c);
        }
        // the read set is loaded once per run, as part of the job (the reference re-reads tmp/_reads.fa per chromosome, AG:1880): raw text to the first
        // GPU, parsed there, ONE broadcast of the packed buffers to the other GPUs — while the host already prepares the first units
        std::vector<string> errors((size_t)units);
        std::vector<char> done((size_t)units, 0);
        int printed = cp;
        struct Progress { std::vector<string>* errors; std::vector<char>* done; int* printed; int units; std::ofstream* wcp; int failed; } pg{&errors, &done, &printed, units, &wcp, -1};
        auto flush_progress = [](Progress& g) {  // progress lines and checkpoints strictly in unit order, as the reference emits them
            while (g.failed < 0 && *g.printed < g.units && (*g.done)[(size_t)*g.printed]) {
                if (!(*g.errors)[(size_t)*g.printed].empty()) { g.failed = *g.printed; return; }   // reported by the main thread after the workers have returned
                cout << endl << "CHROMOSOME " << *g.printed << ": " << endl;
                cout << "(1) Chromosome loaded" << endl << "(2) Contig alignment loaded" << endl << "(3) Read alignment loaded" << endl
                     << "(4) Contigs extended" << endl << "(5) Contigs scaffolded" << endl;
                if (system("ps euf >> mem.txt")) {}
                *g.wcp << *g.printed + 1 << endl;
                (*g.printed)++;
            }
        };
        static void (*flush_fn)(Progress&) = flush_progress;
        auto on_done = [](int unit, int rc, const char* error, void* user) {   // serialised by the library
            Progress& g = *(Progress*)user;
            if (rc != 0) (*g.errors)[(size_t)unit] = error && *error ? error : "UNKNOWN ERROR";
            (*g.done)[(size_t)unit] = 1;
            flush_fn(g);
        };
        int prefetch = 4;
        if (const char* e = getenv("AG_PREFETCH")) prefetch = atoi(e);
        std::vector<int> todo; for (int u = cp; u < units; u++) todo.push_back(u);
        const int job_rc = ag_run_job_files(ctxs.data(), (int)ctxs.size(), "tmp", "tmp/_reads.fa", todo.data(), (int)todo.size(), prefetch, on_done, &pg);
        flush_progress(pg);
        if (pg.failed >= 0) die(errors[(size_t)pg.failed]);   // the reference prints the message and exits at the failing chromosome (exit(-1))
        if (job_rc != 0 && printed < units) die(ag_last_error(ctxs[0]));   // the read set could not be loaded
        if (getenv("AG_STATS")) {
            for (size_t g = 0; g < ctxs.size(); g++) {
                ag_stats s; ag_get_stats(ctxs[g], &s);
                fprintf(stderr, "[ag] gpu %d: parse %.3f s, device section %.3f s, post %.3f s; kernels ms: prep %.2f sort %.2f nodes %.2f finalize %.2f edges %.2f cc %.2f chains %.2f walk %.2f mat %.2f; launches %lu\n",
                        devices[g], s.s_parse, s.s_device_section, s.s_post, s.ms_prep, s.ms_sort, s.ms_nodes, s.ms_finalize, s.ms_edges, s.ms_components, s.ms_chains, s.ms_walk,
                        s.ms_materialize, (unsigned long)s.kernel_launches);
            }
        }
        for (ag_ctx* c : ctxs) ag_destroy(c);
    }

    {
        ag_params p; p.k = o.k; p.insert_variation = o.iv; p.coverage = o.cov; p.device = devices[0];
        ag_ctx* rc = nullptr;
        if (ag_create(&p, &rc) != 0) die(ag_create_error());
        try {
            ag_refinement("tmp", units, genome_ids, contig_ids, o.tagUnique, o.ext, o.rmn, refine_blat, rc, true);  // #define TEST (AG:24) => in.fa / ex.fa
        } catch (const AgHostError& e) { ag_destroy(rc); die(e.msg); }
        ag_destroy(rc);
    }
    if (o.tagMis == 1) {
        // removeMisassembly (AG:4281-4297) on both output files: aligner command lines verbatim (makeAlignment, AG:3821-3850), coverage pile-up on the GPU
        struct Mis { int low, high; } mis{o.low, o.high};
        auto align = [](const char* id_, void* user) -> int {
            const Mis& m = *(const Mis*)user; const string id = id_;
            std::stringstream lo, hi; lo << m.low; hi << m.high;
            string c = "bowtie2-build -f tmp/_" + id + "_contigs.fa tmp/_" + id + "_contigs > bowtie_doc.txt 2> bowtie_doc.txt";
            if (system(c.c_str())) {}
            c = "bowtie2 -f --no-mixed -k 1 -p 8 -I " + lo.str() + " -X " + hi.str() + " --no-discordant -x tmp/_" + id + "_contigs -1 tmp/_reads_1.fa -2 tmp/_reads_2.fa --reorder > tmp/_reads_" + id +
                "_contigs.bowtie 2> bowtie_doc.txt";
            if (system(c.c_str())) {}
            return blat_call("tmp/_genome.fa", "tmp/_" + id + "_contigs.fa", "tmp/_" + id + "_contigs_genome.psl") ? 1 : 0;
        };
        ag_params p; p.k = o.k; p.insert_variation = o.iv; p.coverage = o.cov; p.device = devices[0];
        ag_ctx* c = nullptr;
        if (ag_create(&p, &c) != 0) die(ag_create_error());
        if (ag_remove_misassembly_file(c, o.ext.c_str(), "extended", o.cov, "tmp", align, &mis) != 0) die(ag_last_error(c));
        if (ag_remove_misassembly_file(c, o.rmn.c_str(), "remaining", o.cov, "tmp", align, &mis) != 0) die(ag_last_error(c));
        ag_destroy(c);
        cout << endl << "(6) Misassemblies removed" << endl;
    }
    time_t end = time(NULL);
    cout << endl << "FINISHED SUCCESSFULLY for " << end - start << " seconds (" << endAlign - startAlign << " seconds for alignment) :-)" << endl;
    return 0;
}

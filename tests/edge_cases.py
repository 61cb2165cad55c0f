"""Edge-case work directories shared by the CPU (emulation) and GPU tests: empty / degenerate inputs the reference accepts."""
import os


def _tmp(work, name):
    return os.path.join(work, "tmp", name)


def make(harness, work, kind):
    """Generate a small case and mutate its tmp/ inputs in place."""
    harness.synth(work, genome_bp=20000, coverage=40, seed=61, contig_len=3000, contig_gap=400)
    if kind == "empty_sam":            # no read alignment at all: the unit is built from the contigs only
        open(_tmp(work, "_reads_genome.0.bowtie"), "w").close()
    elif kind == "only_unaligned":     # every record is an unaligned pair
        lines = open(_tmp(work, "_reads_genome.0.bowtie")).read().split("\n")
        with open(_tmp(work, "_reads_genome.0.bowtie"), "w") as f:
            for l in lines:
                if l:
                    x = l.split("\t")
                    f.write(f"{x[0]}\t{77 if int(x[1]) in (99, 83) else 141}\t*\t0\t0\t*\t*\t0\t0\t*\t*\tYT:Z:UP\n")
    elif kind == "empty_psl":          # no contig alignment: no contiMers, _initial_contigs is empty
        open(_tmp(work, "_contigs_genome.0.psl"), "w").close()
    elif kind == "empty_both":
        open(_tmp(work, "_reads_genome.0.bowtie"), "w").close()
        open(_tmp(work, "_contigs_genome.0.psl"), "w").close()
    elif kind == "k_equals_readlen":   # --kMer == read length: the offset loop of AlignGraph.cpp:1681 has zero iterations
        p = _tmp(work, "_command.txt")
        t = open(p).read().split("\n")
        t[t.index("--kMer") + 1] = "100"
        open(p, "w").write("\n".join(t))
    elif kind == "k_readlen_minus_1":  # exactly one call per alignment
        p = _tmp(work, "_command.txt")
        t = open(p).read().split("\n")
        t[t.index("--kMer") + 1] = "99"
        open(p, "w").write("\n".join(t))
    elif kind == "coverage_zero":      # nothing is filtered
        p = _tmp(work, "_command.txt")
        t = open(p).read().split("\n")
        t[t.index("--coverage") + 1] = "0"
        open(p, "w").write("\n".join(t))
    elif kind == "coverage_huge":      # everything outside contigs is filtered
        p = _tmp(work, "_command.txt")
        t = open(p).read().split("\n")
        t[t.index("--coverage") + 1] = "100000"
        open(p, "w").write("\n".join(t))
    elif kind == "all_n_reads":        # reads made of 'N' only: every base count lands in the N counter (AlignGraph.cpp:1349)
        lines = open(_tmp(work, "_reads.fa")).read().split("\n")
        with open(_tmp(work, "_reads.fa"), "w") as f:
            for l in lines:
                if l:
                    f.write((l if l.startswith(">") else "N" * len(l)) + "\n")
    elif kind == "lowercase_reads":    # lower-case bases are not ACGT for the reference: counted as N, copied verbatim into tails
        lines = open(_tmp(work, "_reads.fa")).read().split("\n")
        with open(_tmp(work, "_reads.fa"), "w") as f:
            for i, l in enumerate(lines):
                if l:
                    f.write((l if l.startswith(">") or (i // 2) % 3 else l.lower()) + "\n")
    else:
        raise ValueError(kind)


KINDS = ["empty_sam", "only_unaligned", "empty_psl", "empty_both", "k_equals_readlen", "k_readlen_minus_1", "coverage_zero", "coverage_huge",
         "all_n_reads", "lowercase_reads"]

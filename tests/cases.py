"""Synthetic cases shared by the test-suite, the golden-fixture script and the smoke test (agsynth options)."""

# Small cases whose REAL-reference outputs are committed under tests/golden/<name>/ (made by tests/golden/make_golden.py).
GOLDEN = {
    "plain": dict(genome_bp=30000, coverage=50, contig_len=4000, contig_gap=600, seed=11),
    "mix": dict(genome_bp=30000, coverage=40, indel=0.004, softclip=0.3, multi=0.2, unaligned=0.02, lowqual=0.03, n_rate=0.003,
                read_err=0.005, header=1, contig_len=3000, contig_gap=400, seed=12),
    "k7_150_2chr": dict(genome_bp=40000, chroms=2, coverage=40, readlen=150, kmer=7, indel=0.002, softclip=0.2, contig_len=5000, seed=13),
    "two_chr": dict(genome_bp=40000, chroms=2, coverage=45, contig_len=3500, contig_gap=500, softclip=0.2, multi=0.1, seed=15),
    # overlapping contig tiles: positions with two contiMers (several candidates per touch, AlignGraph.cpp:1369-1477) plus indels / multi-hits
    "overlap_ctg": dict(genome_bp=30000, coverage=40, indel=0.002, softclip=0.2, multi=0.1, contig_len=3000, contig_gap=-1200, seed=16),
    "part2": dict(genome_bp=30000, part=2, coverage=30, indel=0.001, insert_sd=80, cov=10, seed=14),
}

# Larger / nastier cases checked live against the oracle (and the reference where it is available).
LIVE = {
    "a1": dict(genome_bp=100000, coverage=40, indel=0.002, softclip=0.2, multi=0.1, unaligned=0.02, n_rate=0.002, read_err=0.005, header=1,
               contig_len=5000, seed=1),
    "a4": dict(genome_bp=60000, coverage=60, indel=0.01, softclip=0.5, multi=0.5, seed=4, kmer=3, readlen=60, insert_mean=300, contig_len=2000),
    "deep": dict(genome_bp=20000, coverage=400, insert_sd=120, contig_len=3000, seed=5, cov=30),
    "lowcov": dict(genome_bp=50000, coverage=8, seed=6, cov=3, contig_len=8000),
    # BASELINE configs[3] / [4] shapes at test size: 8 chromosomes 2x150 k=7; 2 chromosomes --part 4 (8 units)
    "c4_shape": dict(genome_bp=400000, chroms=8, coverage=50, readlen=150, kmer=7, insert_mean=500, insert_sd=50, seed=31),
    "c5_shape": dict(genome_bp=200000, chroms=2, part=4, coverage=50, readlen=150, kmer=7, seed=32, indel=0.001),
    # a 160 kbp contig: the walk emits a > 100 kbp contig, which switches on the reference's 1000-position scan skip (AlignGraph.cpp:2194-2202)
    "longcontig": dict(genome_bp=500000, coverage=40, contig_len=160000, contig_gap=3000, seed=41),
    # contigs longer than LARGE_CHUNK = 1,000,000 bp are cut into chunks (AlignGraph.cpp:3277-3293); 1.28 Mbp walks, skip rule active
    "bigchunk": dict(genome_bp=3000000, coverage=12, contig_len=1300000, contig_gap=20000, seed=51, cov=4),
    # overlapping contig tiles: positions holding two contiMers (several candidates per touch, AlignGraph.cpp:1369-1477; nodes of such
    # positions live in the overflow pool; every edge around them goes through the generic edge sweep)
    "overlap": dict(genome_bp=60000, coverage=40, indel=0.002, softclip=0.2, multi=0.1, contig_len=3000, contig_gap=-1200, seed=61),
    # very wide insert distribution at 1200x: up to 75 nodes per position, i.e. successor items beyond the 32-bit item mask of the node
    # sweep (tiles flagged for the full generic edge sweep), long overflow chains
    "manyitems": dict(genome_bp=16000, coverage=1200, insert_mean=5000, insert_sd=2500, contig_len=3000, seed=62, cov=3),
    "nocontigs": dict(genome_bp=30000, coverage=50, seed=7, contig_len=150, contig_gap=5000),
}


def add_misassemblies(work_dir, seed=7):
    """Append three contigs that removeMisassembly (AlignGraph.cpp:3821-4297) has to act on to <work_dir>/contigs.fa — none of them is in the
    contig alignment, so they pass through the hot loop into remainingContigs.fa untouched: (1) a chimera of two genome stretches around
    700 random bases (no read covers the junk: the contig is broken into two parts), (2) a genome stretch with 350 random bases at its end
    (the tail is cut off), (3) 320 random bases (removed altogether)."""
    import os
    import random
    rnd = random.Random(seed)
    g = "".join(l.strip() for l in open(os.path.join(work_dir, "genome.fa")) if not l.startswith(">"))
    junk = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    extra = [("chimera", g[2000:3500] + junk(700) + g[12000:13500]), ("junktail", g[20000:21800] + junk(350)), ("alljunk", junk(320))]
    with open(os.path.join(work_dir, "contigs.fa"), "a") as f:
        for name, seq in extra:
            f.write(">" + name + "\n")
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")

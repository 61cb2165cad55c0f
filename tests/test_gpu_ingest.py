"""GPU-side text ingestion (aligngraph_b200/csrc/ag_ingest.cuh) against the host parsers: the packed reads and the surviving alignment
tuples produced by the device kernels from the raw bytes of tmp/_reads.fa / tmp/_reads_genome.N.bowtie must be IDENTICAL, array for
array, to what ag_parse_reads / ag_parse_sam (checked against the oracle and the reference by the CPU suite) produce — on every
golden / live / edge case, on the 1,000,000-id batch boundary (AlignGraph.cpp:1259), and malformed files must take the host parser."""
import ctypes as C
import os
import shutil

import pytest

import cases
import edge_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ag():
    import aligngraph_b200 as m
    m.load_library()
    return m


def _bytes(ptr, n):
    return C.string_at(ptr, n) if ptr and n else b""


def snapshot(ag, harness, work, host_parse):
    """(reads arrays, per-unit (aln bytes, ext bytes), stats) through the device kernels or (host_parse) the host parsers."""
    p = harness.read_command(work)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.set_option("host_parse", 1 if host_parse else 0)
    ctx.load_reads_fasta(os.path.join(work, "tmp", "_reads.fa"))
    b, m, l, n, s2, sm = ctx.get_reads()
    reads = (n, s2, sm, _bytes(b, 2 * n * s2 * 4), _bytes(m, 2 * n * sm * 4), _bytes(l, 2 * n))
    units = []
    for u in range(harness.n_units(work)):
        ctx.prepare_unit(os.path.join(work, "tmp"), u)
        v = ctx.get_unit()
        units.append((v.n_aln, _bytes(v.aln, v.n_aln * 32), v.n_ext, _bytes(v.ext, v.n_ext * 12)))
    st = ctx.stats()
    ctx.close()
    return reads, units, st


def check_same(ag, harness, work, expect_device=True):
    harness.prepare_tmp(work)
    rd, ud, sd = snapshot(ag, harness, work, host_parse=False)
    rh, uh, sh = snapshot(ag, harness, work, host_parse=True)
    assert sh["sam_device"] == 0 and sh["reads_device"] == 0
    assert rd == rh, "packed reads differ"
    assert len(ud) == len(uh)
    for u, (a, b) in enumerate(zip(ud, uh)):
        assert a[0] == b[0], f"unit {u}: {a[0]} alignments from the device, {b[0]} from the host parser"
        assert a == b, f"unit {u}: alignment tuples differ"
    if expect_device:
        assert sd["reads_device"] == 1 and sd["sam_device"] == len(ud) and sd["sam_host"] == 0, sd
    return ud, sd


@pytest.mark.parametrize("name", sorted(cases.GOLDEN))
def test_device_ingest_equals_host_parsers_golden(ag, harness, workdir, name):
    harness.synth(workdir, **cases.GOLDEN[name])
    check_same(ag, harness, workdir)


@pytest.mark.parametrize("name", ["a1", "a4", "deep", "c4_shape", "c5_shape", "overlap", "nocontigs"])
def test_device_ingest_equals_host_parsers_live(ag, harness, workdir, name):
    harness.synth(workdir, **cases.LIVE[name])
    ud, _ = check_same(ag, harness, workdir)
    assert sum(u[0] for u in ud) > 0


@pytest.mark.parametrize("kind", edge_cases.KINDS)
def test_device_ingest_edge_cases(ag, harness, workdir, kind):
    edge_cases.make(harness, workdir, kind)
    check_same(ag, harness, workdir)


@pytest.mark.parametrize("multi", [0.0, 0.4])
def test_device_ingest_batch_boundary(ag, harness, workdir, multi):
    """Read ids starting at 999,900 cross the 1,000,000-id batch boundary: the record lost there (AlignGraph.cpp:1259) and the groups
    it splits must come out of the device kernels exactly as out of the host parser; outputs equal the oracle's."""
    gpu, ora = os.path.join(workdir, "gpu"), os.path.join(workdir, "ora")
    harness.synth(gpu, genome_bp=20000, coverage=40, seed=21, contig_len=3000, multi=multi)
    tmp = os.path.join(gpu, "tmp")
    shift = 999_900
    reads = open(os.path.join(tmp, "_reads.fa")).read().split("\n")
    rl = len(reads[1])
    with open(os.path.join(tmp, "_reads.fa"), "w") as f:
        for i in range(shift):
            f.write(f">{i}\n{'A' * rl}\n>{i}\n{'C' * rl}\n")
        for i in range(0, len(reads) - 1, 2):
            f.write(f">{int(reads[i][1:]) + shift}\n{reads[i + 1]}\n")
    sam = open(os.path.join(tmp, "_reads_genome.0.bowtie")).read().split("\n")
    with open(os.path.join(tmp, "_reads_genome.0.bowtie"), "w") as f:
        for line in sam:
            if line:
                x = line.split("\t")
                x[0] = str(int(x[0]) + shift)
                f.write("\t".join(x) + "\n")
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora)
    ud, _ = check_same(ag, harness, gpu)
    p = harness.read_command(gpu)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.load_reads_fasta(os.path.join(tmp, "_reads.fa"))
    ctx.run_unit(tmp, 0)
    st = ctx.stats()
    ctx.close()
    assert st["sam_device"] == 1 and st["reads_device"] == 1
    assert harness.unit_outputs(gpu, 0) == harness.unit_outputs(ora, 0)


def test_malformed_text_takes_the_host_parser(ag, harness, workdir):
    """An empty line ends the SAM for the reference (AlignGraph.cpp:1247) and a multi-line read record is legal FASTA (AlignGraph.cpp:372-396):
    neither is the layout the kernels handle, so both files must be handed to the sequential host parser — same outputs as the oracle."""
    gpu, ora = os.path.join(workdir, "gpu"), os.path.join(workdir, "ora")
    harness.synth(gpu, genome_bp=20000, coverage=40, seed=22, contig_len=3000)
    sam = os.path.join(gpu, "tmp", "_reads_genome.0.bowtie")
    lines = open(sam).read().split("\n")
    cut = (len(lines) // 2) & ~1
    open(sam, "w").write("\n".join(lines[:cut]) + "\n\n" + "\n".join(lines[cut:]))
    rd = os.path.join(gpu, "tmp", "_reads.fa")
    rl = open(rd).read().split("\n")
    rl[1] = rl[1][:40] + "\n" + rl[1][40:]          # first read spread over two lines
    open(rd, "w").write("\n".join(rl))
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora)
    harness.prepare_tmp(gpu)
    ctx = ag.Context()
    ctx.load_reads_fasta(rd)
    ctx.run_unit(os.path.join(gpu, "tmp"), 0)
    st = ctx.stats()
    ctx.close()
    assert st["sam_host"] == 1 and st["sam_device"] == 0 and st["reads_host"] == 1 and st["reads_device"] == 0
    assert harness.unit_outputs(gpu, 0) == harness.unit_outputs(ora, 0)


def test_device_ingest_full_size_unit(ag, harness, workdir):
    """BASELINE configs[1] at full size (1.15 M pairs: crosses the batch boundary for real): device == host parser, array for array."""
    harness.synth(workdir, genome_bp=4600000, coverage=50, readlen=100, insert_mean=500, insert_sd=50, kmer=5, cov=20, seed=20260927)
    ud, st = check_same(ag, harness, workdir)
    assert ud[0][0] > 1_100_000


def test_job_makes_only_the_referenced_reads_resident(ag, harness, workdir):
    """ag_run_job_files loads the read set for the id window its units' SAM files span (found in the read file by bisection): unit 1 of a
    two-chromosome job only needs the second half of tmp/_reads.fa.  Same files as the reference; a job over both units needs everything."""
    import cases
    from conftest import golden_dir
    harness.synth(workdir, **cases.GOLDEN["two_chr"])
    harness.prepare_tmp(workdir)
    tmp = os.path.join(workdir, "tmp")
    reads_fa = os.path.join(tmp, "_reads.fa")
    p = harness.read_command(workdir)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.run_job(tmp, [1], reads_fa=reads_fa)
    st = ctx.stats()
    assert st["reads_windowed"] == 1 and st["reads_device"] == 1 and st["sam_device"] == 1, st
    assert st["h2d_bytes"] < 0.75 * os.path.getsize(reads_fa) + os.path.getsize(os.path.join(tmp, "_reads_genome.1.bowtie")) + 200000
    g = golden_dir("two_chr")
    for pat in harness.UNIT_FILES:
        assert open(os.path.join(tmp, pat.format(1)), "rb").read() == open(os.path.join(g, pat.format(1)), "rb").read(), pat
    ctx.reset_stats()
    ctx.run_job(tmp, [0, 1], reads_fa=reads_fa)
    st = ctx.stats()
    assert st["reads_windowed"] == 0 and st["reads_device"] == 1 and st["sam_device"] == 2, st
    for u in (0, 1):
        for pat in harness.UNIT_FILES:
            assert open(os.path.join(tmp, pat.format(u)), "rb").read() == open(os.path.join(g, pat.format(u)), "rb").read(), pat
    ctx.close()


def test_read_window_too_small_loads_the_whole_set(ag, harness, workdir):
    """The window is only an estimate (first / last record of each SAM).  A file that references a read outside it — here a record of the other
    chromosome's id range moved into the middle of unit 1's SAM, which also makes the file unsorted — must be noticed on the device, the whole
    read set loaded, and the file handed to the host parser: same outputs as the oracle."""
    import cases
    gpu, ora = os.path.join(workdir, "gpu"), os.path.join(workdir, "ora")
    harness.synth(gpu, **cases.GOLDEN["two_chr"])
    tmp = os.path.join(gpu, "tmp")
    s0 = open(os.path.join(tmp, "_reads_genome.0.bowtie")).read().split("\n")
    s1 = open(os.path.join(tmp, "_reads_genome.1.bowtie")).read().split("\n")
    s1 = [l for l in s1 if l]
    mid = (len(s1) // 2) & ~1
    s1[mid:mid] = s0[10:12]            # a pair of unit 0 (positions are valid in unit 1 as well: same chromosome length)
    open(os.path.join(tmp, "_reads_genome.1.bowtie"), "w").write("\n".join(s1) + "\n")
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora)
    harness.prepare_tmp(gpu)
    p = harness.read_command(gpu)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.run_job(tmp, [1], reads_fa=os.path.join(tmp, "_reads.fa"))
    st = ctx.stats()
    ctx.close()
    assert st["reads_windowed"] == 1 and st["reads_device"] == 2 and st["sam_host"] == 1, st
    assert harness.unit_outputs(gpu, 1) == harness.unit_outputs(ora, 1)


def test_job_over_two_contexts_on_one_gpu(ag, harness, workdir):
    """ag_run_job_files over two contexts that share a GPU (the second one's text staging and host post passes run under the first one's
    kernels; the reads reach it by a device-local copy): same per-unit files as the reference, whichever context ran which unit."""
    import cases
    from conftest import golden_dir
    harness.synth(workdir, **cases.GOLDEN["two_chr"])
    harness.prepare_tmp(workdir)
    tmp = os.path.join(workdir, "tmp")
    p = harness.read_command(workdir)
    ctxs = [ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"]) for _ in range(2)]
    g = golden_dir("two_chr")
    for rep in range(3):
        for u in (0, 1):
            for pat in harness.UNIT_FILES:
                if os.path.exists(os.path.join(tmp, pat.format(u))):
                    os.remove(os.path.join(tmp, pat.format(u)))
        ag.Context.run_job_on(ctxs, tmp, [0, 1], reads_fa=os.path.join(tmp, "_reads.fa"))
        for u in (0, 1):
            for pat in harness.UNIT_FILES:
                assert open(os.path.join(tmp, pat.format(u)), "rb").read() == open(os.path.join(g, pat.format(u)), "rb").read(), (rep, pat, u)
    st = [c.stats() for c in ctxs]
    assert sum(s["sam_device"] for s in st) == 6 and sum(s["sam_host"] for s in st) == 0, st
    for c in ctxs[::-1]:
        c.close()

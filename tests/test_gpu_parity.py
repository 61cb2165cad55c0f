"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the committed reference outputs."""
import os
import shutil

import pytest

import cases
from conftest import compare_with_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ag():
    import aligngraph_b200 as m
    m.load_library()
    return m


def run_cuda(ag, harness, work, dump=False):
    p = harness.read_command(work)
    harness.prepare_tmp(work)  # tmp/_contigs.fa, tmp/_genome.N.fa — input normalisation, outside the hot path
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.load_reads_fasta(os.path.join(work, "tmp", "_reads.fa"))
    if dump:
        ctx.keep_node_counts(True)
    dumps = []
    for u in range(harness.n_units(work)):
        ctx.prepare_unit(os.path.join(work, "tmp"), u)
        ctx.build()
        if dump:
            dumps.append(ctx.dump_nodes_text())
        ctx.extend()
        ctx.write_unit(os.path.join(work, "tmp"), u)
    st = ctx.stats()
    ctx.close()
    return st, dumps


@pytest.mark.parametrize("name", sorted(cases.GOLDEN))
def test_cuda_matches_reference_golden(ag, harness, workdir, name):
    harness.synth(workdir, **cases.GOLDEN[name])
    st, _ = run_cuda(ag, harness, workdir)
    assert st["kernel_launches"] > 0
    compare_with_golden(harness, workdir, name)


@pytest.mark.parametrize("name", sorted(cases.LIVE))
def test_cuda_matches_oracle_nodes_and_files(ag, harness, workdir, name):
    gpu = os.path.join(workdir, "gpu")
    ora = os.path.join(workdir, "ora")
    harness.synth(gpu, **cases.LIVE[name])
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    st, dumps = run_cuda(ag, harness, gpu, dump=True)
    for u in range(harness.n_units(ora)):
        want = open(os.path.join(ora, "tmp", f"_nodes.{u}.txt"), "rb").read()
        assert dumps[u] == want, "node table differs from the oracle"
        assert harness.unit_outputs(gpu, u) == harness.unit_outputs(ora, u)
    # the > 100 kbp contig case must have gone through the exact sequential replay (skip rule), the others must not
    assert st["walk_fallback"] == (1 if name in ("longcontig", "bigchunk") else 0)


def test_run_unit_files_is_the_five_calls(ag, harness, workdir):
    """ag_run_unit_files == loadGenome..scaffoldContigs on the reference's tmp/ contract."""
    harness.synth(workdir, **cases.GOLDEN["plain"])
    harness.prepare_tmp(workdir)
    ctx = ag.Context()
    ctx.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    ctx.run_unit(os.path.join(workdir, "tmp"), 0)
    compare_with_golden(harness, workdir, "plain")


def test_array_level_equals_file_level(ag, harness, workdir):
    """Feeding the staged packed arrays back through ag_begin_unit / ag_set_contimers (explicit contiMer table) or ag_set_contig_threads
    (compact form, table derived on the device) / ag_add_alignments gives the same FASTA."""
    harness.synth(workdir, **cases.GOLDEN["mix"])
    harness.prepare_tmp(workdir)
    p = harness.read_command(workdir)
    a = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    a.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    a.prepare_unit(os.path.join(workdir, "tmp"), 0)
    v = a.get_unit()
    b = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    b.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    b.begin_unit(0, v.ref, v.n_ref)
    b.set_contimers(v.cm_start, v.cm, v.n_cm, v.chain_pos, v.chain_base, v.ref + v.n_ref, v.n_tail)
    b.add_alignments(v.aln, v.n_aln, v.ext, v.n_ext)
    c = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])   # compact form: contig threads
    c.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    c.begin_unit(0, v.ref, v.n_ref)
    c.set_contig_threads(v.threads, v.n_threads, v.chain_pos, v.chain_base, v.n_cm, v.ref + v.n_ref, v.n_tail)
    c.add_alignments(v.aln, v.n_aln, v.ext, v.n_ext)
    b.build(); b.extend()
    c.build(); c.extend()
    a.build(); a.extend()
    assert a.text(1) == b.text(1) and a.text(2) == b.text(2) and len(a.text(1)) > 0
    assert c.text(1) == b.text(1) and c.text(2) == b.text(2)
    g = os.path.join(os.path.dirname(__file__), "golden", "mix")
    assert b.text(1) == open(os.path.join(g, "_pre_extended_contigs.0.fa"), "rb").read()
    assert b.text(2) == open(os.path.join(g, "_extended_contigs.0.fa"), "rb").read()


def test_fused_step_equals_staged_calls(ag, harness, workdir):
    """ag_process (build queued, walk queued behind it, ONE synchronisation) == ag_build + ag_extend (the build checked on its own)."""
    harness.synth(workdir, **cases.GOLDEN["mix"])
    harness.prepare_tmp(workdir)
    p = harness.read_command(workdir)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    ctx.prepare_unit(os.path.join(workdir, "tmp"), 0)
    ctx.build(); ctx.extend()
    staged = (ctx.text(1), ctx.text(2))
    for _ in range(3):
        ctx.process()
        assert (ctx.text(1), ctx.text(2)) == staged
    ctx.set_option("fused_extend", 0)   # emission filter on the host between walk and materialisation (two synchronisations)
    ctx.process()
    assert (ctx.text(1), ctx.text(2)) == staged
    ctx.set_option("fused_extend", 1)
    ctx.set_option("scan_onepass", 0)   # prefix sums by the three-launch scan instead of the single-pass look-back kernel
    ctx.process()
    assert (ctx.text(1), ctx.text(2)) == staged
    g = os.path.join(os.path.dirname(__file__), "golden", "mix")
    assert staged[0] == open(os.path.join(g, "_pre_extended_contigs.0.fa"), "rb").read()
    ctx.close()


def test_full_size_properties(ag, harness, workdir):
    """A larger unit (crosses many tiles; several hundred thousand alignments): size-independent properties — the device build is
    deterministic across runs, every emitted contig is non-empty ACGTN text, headers are strictly ordered by start position."""
    harness.synth(workdir, genome_bp=1000000, coverage=50, seed=99)
    outs = []
    for rep in range(2):
        st, _ = run_cuda(ag, harness, workdir)
        outs.append(harness.unit_outputs(workdir, 0))
    assert outs[0] == outs[1]
    pre = outs[0][1].decode().splitlines()
    starts = [int(l.split(", ")[3]) for l in pre if l.startswith(">")]
    assert starts == sorted(starts) and len(starts) > 100
    assert all(set(l) <= set("ACGTN") for l in pre if not l.startswith(">"))
    assert st["n_nodes"] > 1000000


def test_baseline_config1_full_size_against_oracle(ag, harness, workdir):
    """BASELINE.json configs[1] at FULL size (4.6 Mbp unit, 1.15 M pairs 2x100 at 50x, k=5, coverage=20 — crosses the 1,000,000-pair
    batch boundary of AlignGraph.cpp:1259): every per-unit output byte-identical to the CPU restatement."""
    gpu = os.path.join(workdir, "gpu")
    ora = os.path.join(workdir, "ora")
    harness.synth(gpu, genome_bp=4600000, coverage=50, readlen=100, insert_mean=500, insert_sd=50, kmer=5, cov=20, seed=20260927)
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora)
    st, _ = run_cuda(ag, harness, gpu)
    assert harness.unit_outputs(gpu, 0) == harness.unit_outputs(ora, 0)
    assert st["n_aln"] > 1_100_000 and st["walk_fallback"] == 0


@pytest.mark.parametrize("name", ["mix", "manyitems", "overlap"])
def test_capacity_regrow_paths(ag, harness, workdir, name):
    """Node table, node overflow pool and edge overflow pool start far too small (ag_set_option test hooks): the sweeps must notice, grow
    and repeat — same node table and files as the oracle, and the repeat counter shows that the loops ran."""
    gpu, ora = os.path.join(workdir, "gpu"), os.path.join(workdir, "ora")
    harness.synth(gpu, **(cases.GOLDEN[name] if name in cases.GOLDEN else cases.LIVE[name]))
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    p = harness.read_command(gpu)
    harness.prepare_tmp(gpu)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    for opt, v in (("node_cap", 512), ("ovf_cap", 16), ("eovf_cap", 4), ("key_cap", 64), ("cand_cap", 8), ("hwalk_cap", 4), ("rank_rounds", 1), ("bases_cap", 16)):
        ctx.set_option(opt, v)
    ctx.keep_node_counts(True)
    ctx.load_reads_fasta(os.path.join(gpu, "tmp", "_reads.fa"))
    ctx.prepare_unit(os.path.join(gpu, "tmp"), 0)
    ctx.build()
    dump = ctx.dump_nodes_text()
    ctx.extend()
    ctx.write_unit(os.path.join(gpu, "tmp"), 0)
    st = ctx.stats()
    ctx.close()
    assert st["regrows"] >= 6, st
    assert dump == open(os.path.join(ora, "tmp", "_nodes.0.txt"), "rb").read()
    assert harness.unit_outputs(gpu, 0) == harness.unit_outputs(ora, 0)


import edge_cases


@pytest.mark.parametrize("kind", edge_cases.KINDS)
def test_cuda_edge_cases(ag, harness, workdir, kind):
    gpu = os.path.join(workdir, "gpu")
    ora = os.path.join(workdir, "ora")
    edge_cases.make(harness, gpu, kind)
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    st, dumps = run_cuda(ag, harness, gpu, dump=True)
    assert dumps[0] == open(os.path.join(ora, "tmp", "_nodes.0.txt"), "rb").read()
    assert harness.unit_outputs(gpu, 0) == harness.unit_outputs(ora, 0)


@pytest.mark.parametrize("name", ["two_chr", "k7_150_2chr", "part2"])
def test_pipelined_multi_unit_entry_point(ag, harness, workdir, name):
    """ag_run_units_files: host parsing of later units overlaps the GPU work of earlier ones; same files as the reference."""
    harness.synth(workdir, **cases.GOLDEN[name])
    harness.prepare_tmp(workdir)
    p = harness.read_command(workdir)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    ctx.run_units(os.path.join(workdir, "tmp"), 0, harness.n_units(workdir), prefetch=3)
    ctx.close()
    compare_with_golden(harness, workdir, name)


# ---- one FULL-size unit of BASELINE configs[2], [3] and [4] --------------------------------------------------------------------------
import hashlib
import json

_FULL = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize.json")))


@pytest.mark.parametrize("name", sorted(_FULL))
def test_full_size_unit_matches_oracle_fixture(ag, harness, workdir, name):
    """c3_unit: a 12.5 Mbp chromosome of configs[2] (2x100, k=5); c4_unit: a 25 Mbp chromosome of configs[3] (2x150, k=7; 4.17 M pairs, four
    1,000,000-pair batch boundaries); c5_slice: unit 1 of a 100 Mbp chromosome cut by --part 4 (configs[4]; its reads sit in the middle of a
    16.7 M-pair read file of > 5 GB, ingested in 1 GB segments).  The unit goes through ag_run_unit_files (text in, FASTA out); the three
    files must hash to what the CPU restatement produced from the same generator parameters (tests/golden/make_fullsize.py; the
    restatement is pinned to the unmodified reference at size by tests/test_oracle.py)."""
    c = _FULL[name]
    harness.synth(workdir, **c["params"])
    harness.prepare_tmp(workdir)
    p = harness.read_command(workdir)
    ctx = ag.Context(k=p["kMer"], insert_variation=p["insertVariation"], coverage=p["coverage"])
    ctx.load_reads_fasta(os.path.join(workdir, "tmp", "_reads.fa"))
    u = c["unit"]
    ctx.run_unit(os.path.join(workdir, "tmp"), u)
    st = ctx.stats()
    ctx.close()
    assert st["sam_device"] == 1 and st["reads_device"] == 1, st
    for pat in harness.UNIT_FILES:
        data = open(os.path.join(workdir, "tmp", pat.format(u)), "rb").read()
        want = c["files"][pat.format("N")]
        assert len(data) == want["bytes"], pat
        assert hashlib.sha256(data).hexdigest() == want["sha256"], pat

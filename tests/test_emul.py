"""Host emulation of the CUDA kernels' per-thread logic (tests/emul, built from aligngraph_b200/csrc/ag_core.h) against the oracle:
checks the position-parallel formulation (touch fusion, edges from the final table, component-parallel walk), the host parsers and
the post passes on CPU.  The kernels themselves are exercised by the -m gpu tests."""
import os
import shutil

import pytest

import cases
from conftest import compare_with_golden


@pytest.mark.parametrize("name", sorted(cases.GOLDEN))
def test_emulation_matches_golden(harness, workdir, name):
    harness.synth(workdir, **cases.GOLDEN[name])
    harness.run_emul(workdir)
    compare_with_golden(harness, workdir, name)


@pytest.mark.parametrize("name", sorted(cases.LIVE))
def test_emulation_matches_oracle_nodes_and_files(harness, workdir, name):
    emu = os.path.join(workdir, "emu")
    ora = os.path.join(workdir, "ora")
    harness.synth(emu, **cases.LIVE[name])
    shutil.copytree(emu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    log = harness.run_emul(emu, dump_nodes=True)
    n = harness.n_units(ora)
    for u in range(n):
        assert harness.unit_outputs(emu, u) == harness.unit_outputs(ora, u)
        if name in ("longcontig", "bigchunk"):
            assert "sequential walk" in log
        a = open(os.path.join(emu, "tmp", f"_nodes.{u}.txt"), "rb").read()
        b = open(os.path.join(ora, "tmp", f"_nodes.{u}.txt"), "rb").read()
        assert a == b, "node tables differ"


def test_emulation_sequential_walk_path(harness, workdir):
    """The exact sequential replay used when the reference's 1000-position skip (AlignGraph.cpp:2194-2202) applies."""
    harness.synth(workdir, **cases.GOLDEN["mix"])
    log = harness.run_emul(workdir, env={"AG_EMUL_FORCE_SEQUENTIAL": "1"})
    assert "sequential walk" in log
    compare_with_golden(harness, workdir, "mix")


def test_emulation_generic_path_only(harness, workdir):
    """Same result when every alignment goes through the general multi-segment locator / nested candidate enumeration."""
    harness.synth(workdir, **cases.GOLDEN["mix"])
    harness.run_emul(workdir, env={"AG_EMUL_FORCE_GENERIC": "1"})
    compare_with_golden(harness, workdir, "mix")


def test_emulation_generic_edge_sweep_everywhere(harness, workdir):
    """Edges settled inside the node sweep (successor-item bits) and edges from the generic sweep agree: running the generic sweep on
    every tile, for every alignment, must change nothing."""
    emu = os.path.join(workdir, "emu")
    ora = os.path.join(workdir, "ora")
    harness.synth(emu, **cases.GOLDEN["mix"])
    shutil.copytree(emu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    harness.run_emul(emu, dump_nodes=True, env={"AG_EMUL_ALL_EDGES": "1"})
    assert open(os.path.join(emu, "tmp", "_nodes.0.txt"), "rb").read() == open(os.path.join(ora, "tmp", "_nodes.0.txt"), "rb").read()


def test_emulation_clean_input_needs_no_generic_edge_sweep(harness, workdir):
    """All-M CIGARs and non-overlapping contigs: every edge is settled inside the node sweep, no tile is flagged."""
    harness.synth(workdir, **cases.GOLDEN["plain"])
    log = harness.run_emul(workdir)
    assert "flagged_tiles=0/" in log
    compare_with_golden(harness, workdir, "plain")


@pytest.mark.parametrize("name", ["mix", "two_chr", "k7_150_2chr"])
def test_parallel_text_parsers_match_golden(harness, workdir, name):
    """The multi-threaded read / SAM parsers (used for large files) forced onto the small golden cases: '@' header, -k records,
    unaligned records, soft clips, indels."""
    harness.synth(workdir, **cases.GOLDEN[name])
    harness.run_emul(workdir, env={"AG_PARSE_PARALLEL_MIN": "0", "AG_THREADS": "5"})
    compare_with_golden(harness, workdir, name)


@pytest.mark.parametrize("threads", ["1", "7"])
def test_post_passes_do_not_depend_on_the_thread_team(harness, workdir, threads):
    """The FASTA text is formatted by a host thread team (contigs split into byte-balanced chunks); one thread or seven, the files
    are those of the reference."""
    harness.synth(workdir, **cases.GOLDEN["two_chr"])
    harness.run_emul(workdir, env={"AG_THREADS": threads})
    compare_with_golden(harness, workdir, "two_chr")


import edge_cases


@pytest.mark.parametrize("kind", edge_cases.KINDS)
def test_emulation_edge_cases(harness, workdir, kind):
    emu = os.path.join(workdir, "emu")
    ora = os.path.join(workdir, "ora")
    edge_cases.make(harness, emu, kind)
    shutil.copytree(emu, ora)
    harness.run_oracle(ora, dump_nodes=True)
    harness.run_emul(emu, dump_nodes=True)
    assert harness.unit_outputs(emu, 0) == harness.unit_outputs(ora, 0)
    assert open(os.path.join(emu, "tmp", "_nodes.0.txt"), "rb").read() == open(os.path.join(ora, "tmp", "_nodes.0.txt"), "rb").read()


@pytest.mark.parametrize("name", ["mix", "overlap_ctg", "part2"])
def test_per_base_contig_threading_still_matches_golden(harness, workdir, name):
    """The default host path threads contigs in run space (ag_thread_contigs_runs: interval arithmetic on the PSL blocks, chain arrays
    expanded from descriptors); the per-base formulation it replaced stays as the path for inputs the run form does not model.  Both must
    give the reference's files (the default is what every other emulation test runs)."""
    harness.synth(workdir, **cases.GOLDEN[name])
    harness.run_emul(workdir, env={"AG_CONTIG_PERBASE": "1"})
    compare_with_golden(harness, workdir, name)

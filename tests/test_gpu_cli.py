"""The drop-in command line (aligngraph_b200/bin/AlignGraph) against the reference's FINAL outputs: --extendedContig /
--remainingContig FASTA after refinement, via --resume and via a fresh run with the stub aligners on $PATH."""
import os
import subprocess

import pytest

import cases
from conftest import golden_dir

pytestmark = pytest.mark.gpu

CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aligngraph_b200", "bin", "AlignGraph")


def _same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("name", ["plain", "mix", "k7_150_2chr", "two_chr"])
def test_cli_resume_final_fasta(harness, workdir, name):
    harness.synth(workdir, **cases.GOLDEN[name])
    r = subprocess.run([CLI, "--resume"], cwd=workdir, env=harness.stub_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:]
    assert "RESUMED SUCCESSFULLY :-)" in r.stdout and "(5) Contigs scaffolded" in r.stdout and "FINISHED SUCCESSFULLY" in r.stdout
    g = golden_dir(name)
    for f in ("extendedContigs.fa", "remainingContigs.fa"):
        assert _same(os.path.join(workdir, f), os.path.join(g, f)), f
    for u in range(harness.n_units(workdir)):
        for pat in harness.UNIT_FILES:
            assert _same(os.path.join(workdir, "tmp", pat.format(u)), os.path.join(g, pat.format(u)))
    # checkpoint file: "0" from the generator, then one line per finished unit (AlignGraph.cpp:4782)
    cp = open(os.path.join(workdir, "tmp", "_checkpoint.txt")).read().split()
    assert cp == [str(i) for i in range(harness.n_units(workdir) + 1)]


@pytest.mark.parametrize("name", ["plain", "two_chr"])
def test_cli_fresh_run_final_fasta(harness, workdir, name):
    harness.synth(workdir, **cases.GOLDEN[name])
    args = harness.prepare_fresh(workdir)
    rc, out = harness.run_fresh(CLI, workdir, args, timeout=600)
    assert rc == 0, out[-500:]
    assert "(0) Alignment finished" in out
    g = golden_dir(name + "_fresh")
    for f in ("extendedContigs.fa", "remainingContigs.fa"):
        assert _same(os.path.join(workdir, f), os.path.join(g, f)), f
    for f in ("_reads.fa", "_reads_1.fa", "_contigs.fa", "_genome.fa", "_command.txt"):
        assert os.path.getsize(os.path.join(workdir, "tmp", f)) > 0


def test_cli_multi_device_env_same_output(harness, workdir):
    """AG_DEVICES farms units over contexts; with one physical GPU two contexts on device 0 must give the same files."""
    harness.synth(workdir, **cases.GOLDEN["two_chr"])
    env = harness.stub_env(); env["AG_DEVICES"] = "0,0"
    r = subprocess.run([CLI, "--resume"], cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:]
    g = golden_dir("two_chr")
    for f in ("extendedContigs.fa", "remainingContigs.fa"):
        assert _same(os.path.join(workdir, f), os.path.join(g, f)), f
    assert r.stdout.index("CHROMOSOME 0") < r.stdout.index("CHROMOSOME 1")


def test_cli_multi_device_non_acgt_reads(harness, workdir):
    """Units that run on a SECONDARY context must print the original non-ACGT characters of read tails (AlignGraph.cpp:2167), i.e. the
    contexts of a run share one read set including its exception list (ag_broadcast_reads).  Two units, lower-case and 'N' reads,
    contexts 0,0: per-unit files identical to the oracle's, whichever context ran the unit."""
    import shutil
    gpu, ora = os.path.join(workdir, "gpu"), os.path.join(workdir, "ora")
    harness.synth(gpu, genome_bp=40000, chroms=2, coverage=45, contig_len=3500, contig_gap=500, n_rate=0.01, seed=77)
    p = os.path.join(gpu, "tmp", "_reads.fa")
    lines = open(p).read().split("\n")
    with open(p, "w") as f:
        for i, l in enumerate(lines):
            if l:
                f.write((l if l.startswith(">") or (i // 2) % 3 else l.lower()) + "\n")
    shutil.copytree(gpu, ora)
    harness.run_oracle(ora)
    env = harness.stub_env(); env["AG_DEVICES"] = "0,0"; env["AG_PREFETCH"] = "2"
    r = subprocess.run([CLI, "--resume"], cwd=gpu, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:]
    pre = b""
    for u in range(2):
        assert harness.unit_outputs(gpu, u) == harness.unit_outputs(ora, u), f"unit {u}"
        pre += harness.unit_outputs(gpu, u)[1]
    assert any(c in pre for c in (b"a", b"c", b"g", b"t")), "the case must put lower-case tail characters into the output"


def test_cli_usage_and_errors(workdir):
    r = subprocess.run([CLI], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0 and "--read1 is the the first pair" in r.stdout
    r = subprocess.run([CLI, "--bogus"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 255 and "Inputs:" in r.stdout
    r = subprocess.run([CLI, "--resume"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 255 and "CANNOT OPEN FILE!" in r.stdout

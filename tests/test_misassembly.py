"""removeMisassembly (AlignGraph.cpp:3821-4297, --misassemblyRemoval): the host logic (formalize -> aligners -> coverage pile-up -> keep / break /
drop -> corrected_<file>) against the reference's own outputs.  CPU: through the emulation binary with the sequential pile-up, against the
committed golden (tests/golden/misasm_fresh, made by the reference with the stub aligners) and, where the reference is built, a live run.
GPU (-m gpu): the drop-in CLI end to end, the pile-up on the device."""
import os
import shutil
import subprocess

import pytest

import cases
from conftest import golden_dir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "aligngraph_b200", "bin", "AlignGraph")
FINAL = ("extendedContigs.fa", "remainingContigs.fa", "corrected_extendedContigs.fa", "corrected_remainingContigs.fa")


def make_case(harness, d):
    params = dict(cases.GOLDEN["plain"]); params["misasm"] = 1
    harness.synth(d, **params)
    cases.add_misassemblies(d)


def test_misassembly_host_logic_matches_golden(harness, workdir):
    """From the reference's final FASTA (golden) + the user inputs: formalize, stub aligners with the reference's command lines, pile-up, break / drop.
    The chimeric contig must come out as two parts, the all-junk contig must be gone — byte-identical to the reference's corrected_* files."""
    make_case(harness, workdir)
    g = golden_dir("misasm_fresh")
    for f in FINAL[:2]:
        shutil.copy(os.path.join(g, f), workdir)
    for f, idn in (("extendedContigs.fa", "extended"), ("remainingContigs.fa", "remaining")):
        r = subprocess.run([harness.EMUL, "--remove-misassembly", f, idn, "20", "1"], cwd=workdir, env=harness.stub_env(), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    for f in FINAL[2:]:
        assert open(os.path.join(workdir, f), "rb").read() == open(os.path.join(g, f), "rb").read(), f
    rem = open(os.path.join(workdir, "corrected_remainingContigs.fa")).read()
    assert ">chimera : part0\n" in rem and ">chimera : part1\n" in rem and ">alljunk" not in rem


def test_misassembly_host_logic_matches_live_reference(harness, workdir):
    if not harness.have_reference():
        pytest.skip("reference not built here (oracle/_ref absent)")
    ref, ours = os.path.join(workdir, "ref"), os.path.join(workdir, "ours")
    make_case(harness, ref)
    args = harness.prepare_fresh(ref)
    rc, out = harness.run_fresh(os.path.join(harness.REF, "AlignGraph_shipped"), ref, args)
    assert rc == 0 and "(6) Misassemblies removed" in out
    shutil.copytree(ref, ours)
    for f in FINAL[2:]:
        os.remove(os.path.join(ours, f))
    for f, idn in (("extendedContigs.fa", "extended"), ("remainingContigs.fa", "remaining")):   # on the aligner outputs the reference left in tmp/
        r = subprocess.run([harness.EMUL, "--remove-misassembly", f, idn, "20", "0"], cwd=ours, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    for f in FINAL[2:]:
        assert open(os.path.join(ours, f), "rb").read() == open(os.path.join(ref, f), "rb").read(), f
    g = golden_dir("misasm_fresh")
    for f in FINAL:
        assert open(os.path.join(ref, f), "rb").read() == open(os.path.join(g, f), "rb").read(), "golden is stale: " + f


@pytest.mark.gpu
def test_cli_fresh_run_with_misassembly_removal(harness, workdir):
    """The drop-in binary, fresh run with --misassemblyRemoval: all four output files as the reference writes them; the coverage pile-up ran
    on the GPU (the CLI prints the reference's step line)."""
    make_case(harness, workdir)
    args = harness.prepare_fresh(workdir)
    rc, out = harness.run_fresh(CLI, workdir, args, timeout=600)
    assert rc == 0, out[-600:]
    assert "(6) Misassemblies removed" in out and "FINISHED SUCCESSFULLY" in out
    g = golden_dir("misasm_fresh")
    for f in FINAL:
        assert open(os.path.join(workdir, f), "rb").read() == open(os.path.join(g, f), "rb").read(), f


@pytest.mark.gpu
def test_device_pileup_equals_host_pileup(harness, workdir):
    """The coverage pile-up kernel against the sequential host version on the read-vs-contig SAM of the case above (through the C ABI twice:
    device parsers on / off)."""
    import aligngraph_b200 as ag
    make_case(harness, workdir)
    g = golden_dir("misasm_fresh")
    outs = []
    for host in (0, 1):
        d = os.path.join(workdir, f"run{host}")
        shutil.copytree(workdir, d, ignore=shutil.ignore_patterns("run*"))
        for f in FINAL[:2]:
            shutil.copy(os.path.join(g, f), d)
        r = subprocess.run([harness.EMUL, "--remove-misassembly", "remainingContigs.fa", "remaining", "20", "1"], cwd=d, env=harness.stub_env(), capture_output=True, text=True)
        assert r.returncode == 0   # leaves the aligner outputs in tmp/
        os.remove(os.path.join(d, "corrected_remainingContigs.fa"))
        ctx = ag.Context()
        ctx.set_option("host_parse", host)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            rc = ctx._lib.ag_remove_misassembly_file(ctx._h, b"remainingContigs.fa", b"remaining", 20, b"tmp", None, None)
        finally:
            os.chdir(cwd)
        assert rc == 0
        ctx.close()
        outs.append(open(os.path.join(d, "corrected_remainingContigs.fa"), "rb").read())
    assert outs[0] == outs[1] == open(os.path.join(g, "corrected_remainingContigs.fa"), "rb").read()

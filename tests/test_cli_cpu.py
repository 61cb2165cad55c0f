"""CLI behaviour that needs no GPU: usage text, flag validation and error convention (message on stdout + exit status 255)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "aligngraph_b200", "bin", "AlignGraph")


@pytest.fixture(scope="module", autouse=True)
def _build():
    from aligngraph_b200 import build
    build.build()


def run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True)


def test_usage_matches_reference_text(tmp_path):
    r = run([], tmp_path)
    assert r.returncode == 0
    assert r.stdout.startswith("AlignGraph: algorithm for secondary de novo genome assembly guided by closely related references\n"
                               "By Ergude Bao, CS Department, UC-Riverside. All Rights Reserved\n\nAlignGraph --read1 reads_1.fa")
    assert "--covereage coverage" in r.stdout  # the reference's own typo is part of the surface
    assert (tmp_path / "command.txt").exists()


def test_unknown_flag_and_bad_integer(tmp_path):
    assert run(["--nope"], tmp_path).returncode == 255
    (tmp_path / "a.fa").write_text(">0\nACGT\n")
    r = run(["--read1", "a.fa", "--kMer", "5x"], tmp_path)
    assert r.returncode == 255 and "Inputs:" in r.stdout
    r = run(["--read1", "missing.fa"], tmp_path)
    assert r.returncode == 255 and "CANNOT OPEN FILE!" in r.stdout


def test_resume_must_be_alone_and_needs_checkpoint(tmp_path):
    assert run(["--resume", "--kMer", "5"], tmp_path).returncode == 255
    r = run(["--resume"], tmp_path)
    assert r.returncode == 255 and "CANNOT OPEN FILE!" in r.stdout


def test_no_gpu_is_a_loud_failure(tmp_path, harness):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    harness.synth(str(tmp_path), **cases.GOLDEN["plain"])
    r = subprocess.run([CLI, "--resume"], cwd=tmp_path, env=harness.stub_env(), capture_output=True, text=True)
    assert r.returncode == 255 and "no CUDA device" in r.stdout


def _ratio_case(harness, d):
    """`mix` (unaligned pairs, pairs failing the 0.6 filter, multi-hits) with --ratioCheck appended to the resumed command line."""
    import cases
    import shutil
    harness.synth(d, **cases.GOLDEN["mix"])
    shutil.copy(os.path.join(d, "reads_1.fa"), os.path.join(d, "tmp", "_reads_1.fa"))   # checkRatio only counts its '>' lines (AlignGraph.cpp:3762-3768)
    with open(os.path.join(d, "tmp", "_command.txt"), "a") as f:
        f.write("--ratioCheck\n")


RATIO_LINE_MIX = " - 95.2333% reads aligned "   # printed by the unmodified reference (oracle/_ref) on this case; re-checked live below where it exists


def test_ratio_check_line_matches_reference(tmp_path, harness):
    """--ratioCheck (checkRatio, AlignGraph.cpp:3751-3819) prints the reference's percentage, character for character — before the
    per-chromosome loop, so no GPU is needed to see it."""
    d = str(tmp_path / "ours")
    _ratio_case(harness, d)
    r = subprocess.run([CLI, "--resume"], cwd=d, env=harness.stub_env(), capture_output=True, text=True)
    ours = [l for l in r.stdout.splitlines() if "reads aligned" in l]
    assert ours == [RATIO_LINE_MIX], r.stdout[-400:]
    if harness.have_reference():
        d2 = str(tmp_path / "ref")
        _ratio_case(harness, d2)
        rc, out = harness.run_reference(d2)
        assert [l for l in out.splitlines() if "reads aligned" in l] == ours

"""refinement() without BLAT: the built-in containment search (ag_contain_search: exact 24-mer seeds, full-length verification,
local X-drop alignments otherwise) must give refinement the same alignments as the aligner the goldens were made with (the harness's stub
pblat, which both the reference and this repo call), so that the final FASTA is the reference's with NO aligner on $PATH.
CPU: the search's PSL against the stub's, line set for line set, on the database / query files a real run leaves behind.
GPU: the drop-in CLI with a $PATH that has no pblat / blat, final files against the goldens."""
import os
import shutil
import subprocess

import pytest

import cases
from conftest import golden_dir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "aligngraph_b200", "bin", "AlignGraph")


def _truncated_initial_contigs(src, dst):
    """tmp/_short_initial_contigs.N.fa as refinement() writes it (AlignGraph.cpp:2891-2953): contigs cut to 20,000 bases, '>id.size' then."""
    names, seqs = [], []
    for l in open(src):
        l = l.rstrip("\n")
        if l.startswith(">"):
            names.append(l[1:]); seqs.append("")
        elif names:
            seqs[-1] += l
    with open(dst, "w") as f:
        for n, s in zip(names, seqs):
            f.write(f">{int(n)}.{len(s)}\n" if len(s) > 20000 else f">{int(n)}\n")
            s = s[:20000]
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")


@pytest.mark.parametrize("name", ["plain", "mix", "k7_150_2chr", "overlap_ctg"])
def test_builtin_search_equals_stub_aligner(harness, workdir, name):
    g = golden_dir(name)
    for u in range(2):
        db = os.path.join(g, f"_extended_contigs.{u}.fa")
        if not os.path.exists(db):
            break
        q = os.path.join(workdir, f"q{u}.fa")
        _truncated_initial_contigs(os.path.join(g, f"_initial_contigs.{u}.fa"), q)
        a, b = os.path.join(workdir, f"stub{u}.psl"), os.path.join(workdir, f"ours{u}.psl")
        subprocess.run([os.path.join(harness.BIN, "stubs", "pblat"), db, q, "-noHead", a, "-fastMap", "-threads=8"], check=True)
        subprocess.run([harness.EMUL, "--contain-search", db, q, b], check=True)
        la, lb = sorted(open(a).read().splitlines()), sorted(open(b).read().splitlines())
        assert la == lb and (len(la) > 0 or os.path.getsize(q) == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "mix", "k7_150_2chr", "two_chr"])
def test_cli_without_blat_on_path(harness, workdir, name):
    """--resume run with no pblat / blat anywhere on $PATH: the reference would stop with BLAT CALL FAILED!; the built-in search (candidate
    verification on the GPU) lets the run finish with the reference's final FASTA."""
    harness.synth(workdir, **cases.GOLDEN[name])
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:/bin"
    assert shutil.which("pblat", path=env["PATH"]) is None and shutil.which("blat", path=env["PATH"]) is None
    r = subprocess.run([CLI, "--resume"], cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:]
    g = golden_dir(name)
    for f in ("extendedContigs.fa", "remainingContigs.fa"):
        assert open(os.path.join(workdir, f), "rb").read() == open(os.path.join(g, f), "rb").read(), f

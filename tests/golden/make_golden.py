"""Regenerate tests/golden/*: outputs of the UNMODIFIED reference (oracle/_ref, built from /root/reference by `make -C oracle ref`)
on the deterministic synthetic cases in tests/cases.py.  Run in the development container (the reference is absent elsewhere):

    python tests/golden/make_golden.py
"""
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import harness  # noqa: E402
import cases  # noqa: E402

harness.build_tools()
assert harness.have_reference(), "the reference is not available here"
for name, params in cases.GOLDEN.items():
    work = tempfile.mkdtemp(prefix="ag_golden_")
    harness.synth(work, **params)
    rc, out = harness.run_reference(work, optimized=False)  # the README build line (-O0) is the parity oracle
    dst = os.path.join(ROOT, "tests", "golden", name)
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    n = harness.n_units(work)
    for u in range(n):
        for pat in harness.UNIT_FILES:
            shutil.copy(os.path.join(work, "tmp", pat.format(u)), dst)
    for extra in ("extendedContigs.fa", "remainingContigs.fa"):
        if rc == 0 and os.path.exists(os.path.join(work, extra)):
            shutil.copy(os.path.join(work, extra), dst)
    with open(os.path.join(dst, "README.txt"), "w") as f:
        f.write(f"reference exit code {rc}; units {n}; agsynth params {params}\n")
    print(name, "units", n, "reference exit", rc)
    shutil.rmtree(work)

# a FRESH run (no --resume) of the reference AS SHIPPED (prebuilt binary; a g++-13 build traps in task0/task1) with the stub aligners: exercises formalizeInput for reads, distributeAlignments,
# the aligner command lines and refinement; final FASTA only
for name in ("plain", "two_chr"):
    work = tempfile.mkdtemp(prefix="ag_golden_")
    harness.synth(work, **cases.GOLDEN[name])
    args = harness.prepare_fresh(work)
    rc, out = harness.run_fresh(os.path.join(harness.REF, "AlignGraph_shipped"), work, args)
    dst = os.path.join(ROOT, "tests", "golden", name + "_fresh")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    for extra in ("extendedContigs.fa", "remainingContigs.fa"):
        shutil.copy(os.path.join(work, extra), dst)
    n = harness.n_units(work)
    for u in range(n):
        for pat in harness.UNIT_FILES:
            shutil.copy(os.path.join(work, "tmp", pat.format(u)), dst)
    with open(os.path.join(dst, "README.txt"), "w") as f:
        f.write(f"fresh run (stub bowtie2/pblat), reference exit code {rc}; stdout tail: {out[-200:]!r}\n")
    print(name + "_fresh", "reference exit", rc, "ext bytes", os.path.getsize(os.path.join(dst, "extendedContigs.fa")))
    shutil.rmtree(work)

# removeMisassembly (AlignGraph.cpp:3821-4297): a fresh run with --misassemblyRemoval on `plain` plus three contigs it has to act on
# (tests/cases.py add_misassemblies); the stub aligners serve the read-vs-contig and contig-vs-genome alignments as well
work = tempfile.mkdtemp(prefix="ag_golden_")
params = dict(cases.GOLDEN["plain"]); params["misasm"] = 1
harness.synth(work, **params)
cases.add_misassemblies(work)
args = harness.prepare_fresh(work)
rc, out = harness.run_fresh(os.path.join(harness.REF, "AlignGraph_shipped"), work, args)
dst = os.path.join(ROOT, "tests", "golden", "misasm_fresh")
shutil.rmtree(dst, ignore_errors=True)
os.makedirs(dst)
for f in ("extendedContigs.fa", "remainingContigs.fa", "corrected_extendedContigs.fa", "corrected_remainingContigs.fa"):
    shutil.copy(os.path.join(work, f), dst)
with open(os.path.join(dst, "README.txt"), "w") as f:
    f.write(f"fresh run with --misassemblyRemoval (stub bowtie2 / bowtie2_contigs / pblat), reference exit code {rc}; stdout tail: {out[-120:]!r}\n")
print("misasm_fresh reference exit", rc, {f: os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)})
shutil.rmtree(work)

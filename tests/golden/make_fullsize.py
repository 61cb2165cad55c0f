"""sha256 fixtures of the ORACLE's per-unit outputs for ONE full-size unit of each multi-chromosome BASELINE configuration
(tests/golden/fullsize.json), for the -m gpu tests that cannot afford the CPU restatement at that size on the GPU box.

    python tests/golden/make_fullsize.py [name ...]        # development container; minutes to tens of minutes per case

The oracle is pinned to the unmodified reference on every case of tests/cases.py and, at size, by tests/test_oracle.py
(test_oracle_matches_live_reference_full_size_unit).  The generator (tools/agsynth.cpp) is deterministic in its options, so the GPU
box regenerates byte-identical inputs from the parameters stored next to the hashes."""
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize.json")

# one unit of: configs[2] (12.5 Mbp chromosome, 2x100, k=5), configs[3] (25 Mbp chromosome, 2x150, k=7 — 4.17 M pairs: crosses four
# 1,000,000-pair batch boundaries, AlignGraph.cpp:1259 / :390), configs[4] (one --part 4 slice of a 100 Mbp chromosome = 25 Mbp, 2x150, k=7)
CASES = {
    "c3_unit": dict(params=dict(genome_bp=12_500_000, chroms=1, coverage=50, readlen=100, insert_mean=500, insert_sd=50, kmer=5, cov=20, seed=20260925 + 3, user_reads=0), unit=0),
    "c4_unit": dict(params=dict(genome_bp=25_000_000, chroms=1, coverage=50, readlen=150, insert_mean=500, insert_sd=50, kmer=7, cov=20, seed=20260925 + 4, user_reads=0), unit=0),
    "c5_slice": dict(params=dict(genome_bp=100_000_000, chroms=1, part=4, coverage=50, readlen=150, insert_mean=500, insert_sd=50, kmer=7, cov=20, seed=20260925 + 5, user_reads=0), unit=1),
}


def sha_unit(work, u):
    h = {}
    for pat in harness.UNIT_FILES:
        with open(os.path.join(work, "tmp", pat.format(u)), "rb") as f:
            data = f.read()
        h[pat.format("N")] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    return h


def main(names):
    harness.build_tools(with_emul=False)
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names or CASES:
        c = CASES[name]
        work = tempfile.mkdtemp(prefix="ag_full_", dir=os.environ.get("AG_BIG_TMP", "/tmp"))
        t0 = time.time()
        harness.synth(work, **c["params"])
        t1 = time.time()
        harness.run_oracle(work, first=c["unit"], last=c["unit"], capture=False)
        t2 = time.time()
        db[name] = {"params": c["params"], "unit": c["unit"], "files": sha_unit(work, c["unit"]),
                    "made_by": "oracle/ag_oracle.cpp (CPU restatement)", "synth_s": round(t1 - t0, 1), "oracle_s": round(t2 - t1, 1)}
        json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)
        print(name, db[name]["files"], f"synth {t1 - t0:.0f} s, oracle {t2 - t1:.0f} s", flush=True)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main(sys.argv[1:])

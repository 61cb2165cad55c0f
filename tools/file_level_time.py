#!/usr/bin/env python
"""Wall time of the FILE-level path (tmp/ text in -> tmp/ FASTA out) on one GPU, per phase, device ingest vs host parsers.
    python tools/file_level_time.py [--bp 4600000] [--reps 5]"""
import argparse, os, sys, tempfile, time, shutil, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aligngraph_b200 as ag
from tools import synth

ap = argparse.ArgumentParser()
ap.add_argument("--bp", type=int, default=4_600_000)
ap.add_argument("--readlen", type=int, default=100)
ap.add_argument("--kmer", type=int, default=5)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
base = tempfile.mkdtemp(prefix="ag_fl_")
synth.synth(base, genome_bp=a.bp, coverage=50, readlen=a.readlen, insert_mean=500, insert_sd=50, kmer=a.kmer, cov=20, seed=20260927)
tmp = os.path.join(base, "tmp")
out = {}
for mode in ("device", "host"):
    ctx = ag.Context(k=a.kmer, insert_variation=50, coverage=20)
    ctx.set_option("host_parse", 1 if mode == "host" else 0)
    ctx.formalize_inputs(os.path.join(base, "contigs.fa"), os.path.join(base, "genome.fa"), tmp, 1)
    rows = []
    for rep in range(a.reps):
        ctx.reset_stats()
        t0 = time.perf_counter(); ctx.load_reads_fasta(os.path.join(tmp, "_reads.fa")); t1 = time.perf_counter()
        ctx.prepare_unit(tmp, 0); t2 = time.perf_counter()
        t3 = time.perf_counter()
        ctx.process(); t4 = time.perf_counter()
        ctx.write_unit(tmp, 0); t5 = time.perf_counter()
        st = ctx.stats()
        rows.append(dict(reads=t1 - t0, prepare=t2 - t1, build=t3 - t2, extend=t4 - t3, write=t5 - t4, total=t5 - t0,
                         ms_ingest_reads=st["ms_ingest_reads"], ms_ingest_sam=st["ms_ingest_sam"], sam_device=st["sam_device"], reads_device=st["reads_device"]))
    best = min(rows, key=lambda r: r["total"])
    out[mode] = {k: (round(v * 1e3, 2) if isinstance(v, float) and k not in ("ms_ingest_reads", "ms_ingest_sam") else v) for k, v in best.items()}
    ctx.close()
print(json.dumps(out, indent=1))
shutil.rmtree(base, ignore_errors=True)

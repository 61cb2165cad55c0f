#!/usr/bin/env python
"""Full-size run of a BASELINE.json multi-unit configuration through the file-level C ABI on ONE GPU (north-star target: >= 10x the
reference's CPU hot-path time on the 200 Mbp / 8 chromosome / 50x 2x150 / k=7 configuration).

    python tools/scale_run.py --config c4 --json gpurun_out/scale_c4.json      # GPU box
    python tools/scale_run.py --config c4 --oracle-unit 0                      # dev container: sha256 of the oracle's unit-0 outputs

The reference's time for the same configuration is extrapolated from one bounded slice of the same shape timed on the same host
(its hot path is linear in pairs x (L - k): BASELINE.md cost model) — stated in the JSON.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[3]: 200 Mbp, 8 chromosomes, 50x 2x150 (insert 500), k=7, coverage=20
    "c4": dict(genome_bp=200_000_000, chroms=8, coverage=50, readlen=150, insert_mean=500, insert_sd=50, kmer=7, cov=20, seed=20260925 + 4, user_reads=0),
    # BASELINE.json configs[2]: 50 Mbp, 4 chromosomes, 50x 2x100, k=5
    "c3": dict(genome_bp=50_000_000, chroms=4, coverage=50, readlen=100, insert_mean=500, insert_sd=50, kmer=5, cov=20, seed=20260925 + 3, user_reads=0),
    "mini": dict(genome_bp=2_000_000, chroms=2, coverage=50, readlen=150, insert_mean=500, insert_sd=50, kmer=7, cov=20, seed=20260925 + 9, user_reads=0),
}


def sha_unit(work, u):
    h = []
    for pat in ("_initial_contigs.{}.fa", "_pre_extended_contigs.{}.fa", "_extended_contigs.{}.fa"):
        with open(os.path.join(work, "tmp", pat.format(u)), "rb") as f:
            h.append(hashlib.sha256(f.read()).hexdigest()[:16])
    return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--json", default=None)
    ap.add_argument("--oracle-unit", type=int, default=None)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one ag_run_unit_files call per unit, no prefetch")
    ap.add_argument("--prefetch", type=int, default=8)
    ap.add_argument("--ref-slice-bp", type=int, default=2_500_000)
    args = ap.parse_args()
    from tools import synth
    cfg = dict(CONFIGS[args.config])
    work = tempfile.mkdtemp(prefix=f"ag_scale_{args.config}_")
    res = {"config": args.config, "params": cfg}
    try:
        t0 = time.perf_counter()
        meta = synth.synth(work, **cfg)
        res["synth_s"] = round(time.perf_counter() - t0, 1)
        res["pairs"] = meta["pairs"]; res["units"] = meta["units"]
        tmp = os.path.join(work, "tmp")
        if args.oracle_unit is not None:   # dev container: the CPU restatement on one unit, for a checksum comparison
            from oracle import harness
            harness.build_tools(with_emul=False)
            harness.prepare_tmp(work)
            t0 = time.perf_counter()
            harness.run_oracle(work, prepare=False, first=args.oracle_unit, last=args.oracle_unit, capture=False)
            res["oracle_unit_s"] = round(time.perf_counter() - t0, 1)
            res["sha"] = {str(args.oracle_unit): sha_unit(work, args.oracle_unit)}
            print(json.dumps(res))
            return
        import aligngraph_b200 as ag
        from aligngraph_b200 import build
        build.build()
        ctx = ag.Context(k=cfg["kmer"], insert_variation=50, coverage=cfg["cov"], device=0)
        units = ctx.formalize_inputs(os.path.join(work, "contigs.fa"), os.path.join(work, "genome.fa"), tmp, 1)
        t_hot0 = time.perf_counter()
        t0 = time.perf_counter()
        ctx.load_reads_fasta(os.path.join(tmp, "_reads.fa"))
        res["reads_parse_s"] = round(time.perf_counter() - t0, 2)
        per_unit = []
        if args.serial:
            for u in range(units):
                t0 = time.perf_counter()
                ctx.run_unit(tmp, u)
                per_unit.append(round(time.perf_counter() - t0, 3))
        else:   # host parsing of the next units pipelined ahead of the GPU (ag_run_units_files)
            t0 = time.perf_counter()
            ctx.run_units(tmp, 0, units, prefetch=args.prefetch)
            per_unit.append(round(time.perf_counter() - t0, 3))
        t_hot = time.perf_counter() - t_hot0
        st = ctx.stats()
        res.update({"t_hot_s": round(t_hot, 2), "per_unit_s": per_unit, "genome_mbp_per_s": round(cfg["genome_bp"] / 1e6 / t_hot, 3),
                    "read_mbp_per_s": round(2 * cfg["readlen"] * meta["pairs"] / 1e6 / t_hot, 1),
                    "stats": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
                    "sha": {str(u): sha_unit(work, u) for u in range(min(units, 2))}})
        ctx.close()
        # reference on a bounded slice of the same shape, same host
        try:
            from oracle import harness
            harness.build_tools(with_emul=False)
            import bench
            sdir = tempfile.mkdtemp(prefix="ag_scale_ref_")
            scfg = dict(cfg); scfg.update(genome_bp=args.ref_slice_bp, chroms=1, seed=cfg["seed"] + 50)
            d = os.path.join(sdir, "s")
            synth.synth(d, **scfg)
            harness.prepare_tmp(d)
            s, kind = bench.cpu_pass([d])
            res["reference_slice"] = {"bp": args.ref_slice_bp, "hot_s": round(s, 2), "kind": kind, "mbp_per_s": round(args.ref_slice_bp / 1e6 / s, 4),
                                      "extrapolated_full_s": round(s * cfg["genome_bp"] / args.ref_slice_bp, 0)}
            res["speedup_vs_reference_extrapolated"] = round(s * cfg["genome_bp"] / args.ref_slice_bp / t_hot, 1)
            shutil.rmtree(sdir, ignore_errors=True)
        except Exception as e:
            res["reference_slice"] = {"error": str(e)}
        out = json.dumps(res)
        print(out)
        if args.json:
            with open(args.json, "w") as f:
                f.write(out + "\n")
    finally:
        if not args.keep:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

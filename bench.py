#!/usr/bin/env python
"""bench.py — graph-build + contig-extension throughput of the B200 AlignGraph hot path (BASELINE.json `metric`).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N ...]            # the reference's own CPU implementation of the path

Workload (config.workload): --config c2 (default) = BASELINE.json configs[1] — E. coli-sized 4.6 Mbp unit, 50x synthetic 2x100 bp PE reads
(insert 500 +- 50), k = 5, coverage = 20 — ONE such unit per GPU (weak scaling: chromosomes are farmed one per GPU);
--config c3 | c4 | c5 = configs[2..4] (4 x 12.5 Mbp; 8 x 25 Mbp 2x150 k=7; 2 x 100 Mbp --part 4), a fixed job of units dealt round-robin
to the GPUs.  A "step" is one full pass of the hot path over the rank's unit(s).

Printed JSON (one line, rank 0):
  value   Mbp of reference genome per second, all GPUs, inputs already resident in HBM when the timed region starts (device pipeline +
          host post passes, ending with the unit's FASTA text in host memory)
  e2e     the same metric as T_hot of SURVEY.md §8d: every step starts from the tmp/ TEXT files (reads FASTA, genome, PSL, SAM) in host
          memory and ends with the three per-unit FASTA files written — host->device copies of the raw text, device-side parsing, graph
          build, walk, post passes and file output all inside the timed region; this is what the reference arm's number covers too
  roofline  dominant kernel (k_build) against the measured HBM peak;  cpu_baseline  the reference's CPU path on one full-size unit
--impl reference: the unmodified reference (oracle/_ref, else the oracle port) on the SAME configuration, full-size units.
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs.  "c2" (configs[1], the configuration the metric is quoted on) is the default: ONE 4.6 Mbp unit per GPU, weak scaling.
# c3 / c4 / c5 are the multi-chromosome configs[2..4]: a FIXED job of U units, unit u -> rank u mod N (strong scaling).
CONFIGS = {
    "c2": dict(unit_bp=4_600_000, units=None, part=1, readlen=100, kmer=5, seed=20260925 + 2, scaling="weak",
               name="BASELINE configs[1] shape: one 4.6 Mbp unit per GPU"),
    "c3": dict(unit_bp=12_500_000, units=4, part=1, readlen=100, kmer=5, seed=20260925 + 3, scaling="strong",
               name="BASELINE configs[2]: 50 Mbp genome, 4 chromosomes of 12.5 Mbp"),
    "c4": dict(unit_bp=25_000_000, units=8, part=1, readlen=150, kmer=7, seed=20260925 + 4, scaling="strong",
               name="BASELINE configs[3]: 200 Mbp genome, 8 chromosomes of 25 Mbp"),
    "c5": dict(unit_bp=25_000_000, units=8, part=4, readlen=150, kmer=7, seed=20260925 + 5, scaling="strong",
               name="BASELINE configs[4]: 200 Mbp genome, 2 chromosomes of 100 Mbp, --part 4 (8 units of 25 Mbp)"),
}
METRIC = "graph-build+extend Mbp/s"


def shape_of(cfg):
    return dict(coverage=50, readlen=cfg["readlen"], insert_mean=500, insert_sd=50, kmer=cfg["kmer"], cov=20, contig_len=10000, contig_gap=1000, snp=0.01)


def n_units_of(cfg, n_gpus):
    return cfg["units"] if cfg["units"] else n_gpus


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_config(cfg_name, n):
    """The SAME dict for both arms (the driver compares them): it names the workload, not the implementation."""
    cfg = CONFIGS[cfg_name]
    units = n_units_of(cfg, n)
    L = cfg["readlen"]
    return {"workload": f"{cfg['name']}, 50x synthetic 2x{L} bp PE (insert 500+-50), k={cfg['kmer']}, coverage=20, 10 kbp contig tiles, SNP 1%; "
                        f"every step = the whole hot path (loadGenome .. scaffoldContigs, AlignGraph.cpp:4768-4776) over {units} unit(s) from the tmp/ text files",
            "config": cfg_name, "unit_bp": cfg["unit_bp"], "units": units, "pairs_per_unit": int(cfg["unit_bp"] * 50 / (2 * L)), "readlen": L, "k": cfg["kmer"],
            "coverage_threshold": 20, "insert_variation": 50, "part": cfg["part"],
            "parallelism": f"{units} unit(s) over {n} GPU(s), unit u -> GPU u mod N, no data-path collective",
            "l2": "inputs + tables per step (> 1 GB per unit) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_pass(sample_dirs):
    """One pass of the reference CPU implementation over the given single-unit work directories, one process per directory, all
    started together.  Returns (seconds of the hot path = slowest process, kind)."""
    from oracle import harness
    procs = []
    use_ref = harness.have_reference()
    env = dict(os.environ)
    env["PATH"] = os.path.join(harness.BIN, "stubs") + os.pathsep + env["PATH"]
    for d in sample_dirs:
        with open(os.path.join(d, "tmp", "_checkpoint.txt"), "w") as f:
            f.write("0\n")
        if use_ref:   # unmodified reference, -O2, through its --resume door; the hot path is the span of its "(n)" progress lines
            p = subprocess.Popen([os.path.join(harness.REF, "AlignGraph_O2"), "--resume"], cwd=d, env=env, stdout=subprocess.PIPE, text=True)
        else:         # CPU restatement (port)
            p = subprocess.Popen([os.path.join(harness.BIN, "ag_oracle"), "--dir", d, "--no-prepare"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(p)
    spans = []

    def watch(p):
        t_first = t_last = None
        t_start = time.perf_counter()
        for line in p.stdout:
            now = time.perf_counter()
            if line.startswith("CHROMOSOME") and t_first is None:
                t_first = now
            if line.startswith("(5) Contigs scaffolded"):
                t_last = now
        p.wait()
        t_end = time.perf_counter()
        spans.append((t_last - t_first) if (t_first and t_last) else (t_end - t_start))

    th = [threading.Thread(target=watch, args=(p,)) for p in procs]
    [t.start() for t in th]
    [t.join() for t in th]
    return max(spans), ("reference" if use_ref else "port")


def make_unit_dirs(base, cfg_name, n_dirs):
    """n_dirs single-unit work directories of the configuration's unit shape and FULL unit size (one reference process each)."""
    from oracle import harness
    from tools import synth
    cfg = CONFIGS[cfg_name]
    harness.build_tools(with_emul=False)
    dirs = []
    for i in range(n_dirs):
        d = os.path.join(base, f"unit{i}")
        synth.synth(d, genome_bp=cfg["unit_bp"], seed=cfg["seed"] + 100 + i, user_reads=0, **shape_of(cfg))
        harness.prepare_tmp(d)
        dirs.append(d)
    return dirs


def reference_arm(args):
    """The reference's own CPU implementation of the path on the SAME configuration: full-size units, one single-threaded process per
    unit (its hot loop has no threads), min(units, host cores) at a time.  Every step is one full pass over all units."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    cfg = CONFIGS[args.config]
    units = n_units_of(cfg, n)
    cores = min(units, os.cpu_count() or 1)
    base = tempfile.mkdtemp(prefix="ag_bench_ref_")
    try:
        dirs = make_unit_dirs(base, args.config, units)
        kind = "port"
        t_probe0 = time.perf_counter()
        warm = args.warmup
        done_warm = 0
        worst = 0.0
        t0 = time.perf_counter()
        # warm-up passes only warm the page cache for a CPU process; they are capped so that W + K full-size passes fit the driver's limit
        budget_s = float(os.environ.get("AG_REF_BUDGET_S", "1500"))
        for i in range(warm):
            s, kind = cpu_pass(dirs)
            done_warm += 1
            if (time.perf_counter() - t_probe0) + s * (args.steps + (warm - done_warm)) > budget_s:
                break
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s, kind = cpu_pass(dirs)
            worst += s
        wall = time.perf_counter() - t0
        value = units * cfg["unit_bp"] / 1e6 * args.steps / worst
        line = {
            "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "Mbp/s", "n_gpus": n, "steps": args.steps, "warmup": done_warm,
            "ms_per_step": round(1000 * worst / args.steps, 2), "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args.config, n),
            "cpu_baseline": {"value": round(value, 5), "unit": "Mbp/s", "cores": cores, "kind": kind,
                             "sample": f"{units} FULL-size unit(s) of {cfg['unit_bp']} bp, one single-threaded process per unit, {cores} at a time; hot path = span of the "
                                       f"reference's progress lines (CHROMOSOME n .. (5) Contigs scaffolded): text parsing of the tmp/ files included, as for the GPU arm's e2e"},
            "e2e": {"value": round(value, 5), "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(wall, 2),
        }
        emit(line)
    finally:
        shutil.rmtree(base, ignore_errors=True)


# ------------------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------------------------
def bind_rank_cores(world, local):
    """One process per GPU: give every rank its own block of host cores, local to its GPU's NUMA node where NVML knows it."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        mine = None
        try:
            import pynvml
            pynvml.nvmlInit()
            words = (max(allowed) // 64) + 1
            aff = []
            for g in range(world):
                mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), words)
                aff.append(tuple(c for c in allowed if (mask[c // 64] >> (c % 64)) & 1))
            peers = [g for g in range(world) if aff[g] == aff[local]]
            if aff[local] and len(aff[local]) // len(peers) >= 2:
                per = len(aff[local]) // len(peers)
                i = peers.index(local)
                mine = aff[local][i * per:(i + 1) * per]
        except Exception:
            mine = None
        if mine is None:
            per = len(allowed) // world
            if per >= 2:
                mine = allowed[local * per:(local + 1) * per]
        if mine:
            os.sched_setaffinity(0, set(mine))
            os.environ.setdefault("AG_THREADS", str(len(mine)))   # host thread team / parser threads = this rank's core share
    except (AttributeError, OSError):
        pass


def b200_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        bind_rank_cores(world, local)   # before the library starts its thread team
    os.environ.setdefault("AG_MALLOPT", "1")   # this process is the benchmark's own: keep the staging blocks on the heap from step to step
    import aligngraph_b200 as ag
    from aligngraph_b200 import build as _build
    from tools import synth
    if local == 0:
        _build.build()  # no-op when the in-tree library is newer than its sources
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE {world} != --gpus {args.gpus}")
    n = world
    cfg = CONFIGS[args.config]
    units = n_units_of(cfg, n)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic job: rank 0 writes the data set once (user inputs + the tmp/ files Bowtie2 / BLAT would have produced) -----------
    box = [None]
    if rank == 0:
        base = tempfile.mkdtemp(prefix="ag_bench_")
        t0 = time.perf_counter()
        chroms = units // cfg["part"]
        synth.synth(base, genome_bp=cfg["unit_bp"] * units, chroms=chroms, part=cfg["part"], seed=cfg["seed"], user_reads=0, **shape_of(cfg))
        log(f"[bench] synthetic data ({units} unit(s)) in {time.perf_counter() - t0:.1f} s -> {base}")
        box[0] = base
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    base = box[0]
    tmp = os.path.join(base, "tmp")
    reads_fa = os.path.join(tmp, "_reads.fa")
    ctx = ag.Context(k=cfg["kmer"], insert_variation=50, coverage=20, device=local)
    if rank == 0:
        got = ctx.formalize_inputs(os.path.join(base, "contigs.fa"), os.path.join(base, "genome.fa"), tmp, cfg["part"])
        assert got == units, (got, units)
    barrier()
    my_units = [u for u in range(units) if u % n == rank]

    # ---- (1) `value`: inputs resident in HBM.  Per unit: text ingested + arrays uploaded once (untimed), then K timed steps of the device
    # pipeline + host post passes ending with the unit's FASTA text in host memory --------------------------------------------------------
    ctx.load_reads_for_units(reads_fa, tmp, my_units)   # (only the reads this rank's units reference: the read file is shared by all units)
    W = max(args.warmup, 3); K = args.steps
    ms = 0.0
    stats = None
    check = {}
    clocks = ClockSampler(local) if rank == 0 else None
    for u in my_units:
        ctx.prepare_unit(tmp, u)
        for _ in range(W):
            ctx.process()
        check[u] = ctx.text(1)
        ctx.reset_stats()
        barrier() if len(my_units) == 1 else torch.cuda.synchronize()
        ctx.timer_start()
        for _ in range(K):
            ctx.process()   # graph build + walk queued as one step, one host synchronisation, post passes, FASTA text in host memory
        ms += ctx.timer_stop()
        st = ctx.stats()
        stats = st if stats is None else {k: (stats[k] + st[k]) if k != "walk_fallback" else max(stats[k], st[k]) for k in st}
    barrier()

    # ---- (2) `e2e` = T_hot: every step starts from the tmp/ TEXT files in host memory (page cache) and ends with the three
    # per-unit FASTA files written: reads text -> device (parsed there), per unit genome / PSL / SAM text -> device, graph build, walk,
    # post passes, file output.  The reads are ingested once per step and shared by the rank's units, as the product does per run ------
    # A rank with several units runs them over two contexts on its GPU: the second context's text staging and host post passes overlap the
    # first one's kernels (ag_run_job_files takes any number of contexts; the reads go to the first and are copied on the device).
    n_ctx = args.contexts_per_gpu if args.contexts_per_gpu > 0 else (2 if len(my_units) > 1 else 1)
    ctxs = [ctx] + [ag.Context(k=cfg["kmer"], insert_variation=50, coverage=20, device=local) for _ in range(n_ctx - 1)]

    def e2e_step():
        ag.Context.run_job_on(ctxs, tmp, my_units, reads_fa=reads_fa, prefetch=2)   # ag_run_job_files: reads text -> GPU while the first unit's genome / PSL are parsed

    for _ in range(W):
        e2e_step()
    for c in ctxs:
        c.reset_stats()
    barrier()
    ctx.timer_start()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()   # every context's stream drained before the closing event is recorded
    ms_e2e = ctx.timer_stop()
    st_e2e = ctx.stats()
    for c in ctxs[1:]:
        st = c.stats()
        st_e2e = {k: (st_e2e[k] + st[k]) if k != "walk_fallback" else max(st_e2e[k], st[k]) for k in st}
    clock_info = clocks.stop() if clocks else None   # sampled every 100 ms across both timed regions
    barrier()
    for u in my_units:   # the files the timed region wrote are the ones the resident path produced
        with open(os.path.join(tmp, f"_pre_extended_contigs.{u}.fa"), "rb") as f:
            assert f.read() == check[u] and len(check[u]) > 0

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    cnt = torch.tensor([float(stats[k]) for k in ("n_aln", "n_nodes", "n_keys", "kernel_launches")] + [float(st_e2e["h2d_bytes"]), float(st_e2e["d2h_bytes"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    tot_aln, tot_nodes, tot_keys, tot_launch, e2e_h2d, e2e_d2h = cnt.tolist()

    if rank == 0:
        mbp = units * cfg["unit_bp"] / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        # dominant kernel: k_build (node sweep + common-case edges).  Algorithmic bytes of one launch: per tile key the staged
        # record (prepared alignment + the left mate's bases); per position the contiMer summary, reference base and CSR entry; per node the
        # final records written (16 + 16 + 8 + 4).  Rank 0's launches (K per unit).
        L = cfg["readlen"]
        n_launch = K * len(my_units)
        tma = stats.get("ms_build_kernel", 0) > 0 and stats.get("ms_stage", 0) > 0
        rec_bytes = ((12 + (L + 7) // 8 + 3) // 4 * 4) * 4 if tma else (4 + 32 + (L + 3) // 4 + (L + 7) // 8)   # staged record (48-byte header + 4-bit codes) | index + prepared record + packed words
        alg = stats["n_keys"] / max(len(my_units), 1) * rec_bytes + cfg["unit_bp"] * (8 + 1 + 4) + stats["n_nodes"] / max(len(my_units), 1) * 44
        nodes_ms = (stats["ms_build_kernel"] if stats.get("ms_build_kernel", 0) > 0 else stats["ms_nodes"]) / n_launch
        achieved = alg / nodes_ms / 1e6
        traffic = None
        try:   # DRAM bytes of one k_build launch from the committed `ncu --set full` capture of the c2 workload (profiles/README.md)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["traffic"] if args.config == "c2" else None
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": round(mbp * K / (ms / 1000), 3), "unit": "Mbp/s", "n_gpus": n, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args.config, n),
            "e2e": {"contexts_per_gpu": n_ctx, "value": round(mbp * K / (ms_e2e / 1000), 3), "unit": "Mbp/s", "h2d_bytes_per_step": int(e2e_h2d / K), "d2h_bytes_per_step": int(e2e_d2h / K),
                    "ms_per_step": round(ms_e2e / K, 3),
                    "scope": "T_hot from the tmp/ text files (reads FASTA, genome, PSL, SAM) to the three per-unit FASTA files on disk, through ag_run_job_files "
                             "(= ag_load_reads_fasta + ag_run_unit_files per unit, host parsing overlapped); h2d/d2h bytes summed over all GPUs",
                    "rank0_breakdown_ms_per_step": {"reads_ingest": round(st_e2e["ms_ingest_reads"] / K, 3), "sam_ingest": round(st_e2e["ms_ingest_sam"] / K, 3),
                                                    "host_parse_s": round(st_e2e["s_parse"] / K * 1e3, 3), "device_section": round(st_e2e["s_device_section"] / K * 1e3, 3),
                                                    "post_passes": round(st_e2e["s_post"] / K * 1e3, 3)},
                    "text_parsed_on": {"sam_device": int(st_e2e["sam_device"]), "sam_host": int(st_e2e["sam_host"]), "reads_device": int(st_e2e["reads_device"]),
                                       "reads_host": int(st_e2e["reads_host"])}},
            "gpu_launches": int(tot_launch),
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": "k_build_tma" if tma else "k_build", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "alg_bytes_per_launch": int(alg), "ms_per_launch": round(nodes_ms, 4),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         # SURVEY.md §8(d)'s per-pair figure for the bucketing kernel, B_K1 = 2*ceil(L/4) + 32 + 16*(L-k) (it assumes one materialised 16-byte event
                         # per left-mate k-mer; this formulation never writes those, so `achieved` above uses the bytes it really moves): reported beside it
                         "survey_b_k1": (lambda bk1, pairs, t_sweep, t_bucket: {
                             "bytes_per_pair": bk1, "bytes_per_launch": int(bk1 * pairs),
                             "frac_sweep_kernel": round(bk1 * pairs / t_sweep / 1e6 / peak, 4),
                             "frac_prep_sort_stage_sweep": round(bk1 * pairs / t_bucket / 1e6 / peak, 4)})(
                             2 * ((L + 3) // 4) + 32 + 16 * (L - cfg["kmer"]), stats["n_aln"] / n_launch, nodes_ms,
                             (stats["ms_prep"] + stats["ms_sort"] + stats["ms_nodes"]) / n_launch)},
            "device_ms_per_step": {k[3:]: round(stats[k] / K, 4) for k in stats if k.startswith("ms_")},
            "counts": {"alignments": int(tot_aln / K), "nodes": int(tot_nodes), "tile_keys": int(tot_keys), "units": units},
            "host_s_per_step": {"device_section": round(stats["s_device_section"] / K, 4), "post_passes": round(stats["s_post"] / K, 4)},
            "walk_fallback": int(stats["walk_fallback"]),
        }
        if not args.no_cpu and n == 1:
            try:   # the reference's CPU implementation on ONE full-size unit of this configuration, timed once (~30 s for c2)
                sdir = tempfile.mkdtemp(prefix="ag_bench_cpu_")
                dirs = make_unit_dirs(sdir, args.config, 1)
                s, kind = cpu_pass(dirs)
                line["cpu_baseline"] = {"value": round(cfg["unit_bp"] / 1e6 / s, 5), "unit": "Mbp/s", "cores": 1, "kind": kind,
                                        "sample": f"one FULL-size unit ({cfg['unit_bp']} bp) of this configuration, hot path of the reference's CPU implementation "
                                                  f"(text parsing included) timed once: {s:.2f} s"}
                shutil.rmtree(sdir, ignore_errors=True)
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "Mbp/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    barrier()
    for c in ctxs[::-1]:
        c.close()
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints to fd 1 (NCCL's version banner, ...) was diverted to
    stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # C-level and Python-level stdout -> stderr from here on
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default c2 = configs[1], the one the metric is quoted on)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--contexts-per-gpu", type=int, default=0, help="contexts per GPU in the e2e leg (0 = 2 when a rank has several units, else 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()

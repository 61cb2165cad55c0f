#!/usr/bin/env python
"""bench.py — graph-build + contig-extension throughput of the B200 AlignGraph hot path (BASELINE.json `metric`).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N ...]            # the reference's own CPU implementation of the path

Workload (config.workload): BASELINE.json configs[1] — E. coli-sized 4.6 Mbp unit, 50x synthetic 2x100 bp PE reads (insert 500 +- 50),
k = 5, coverage = 20 — ONE such unit per GPU (weak scaling: chromosomes are farmed one per GPU).  A "step" is one
full pass of the hot path over the rank's unit: positional de Bruijn graph build, coverage filter, extension walk, contig
de-dup/join and scaffolding, ending with the unit's FASTA text in host memory.

Printed JSON (one line, rank 0):
  value   Mbp of reference genome processed per second, all GPUs, inputs already resident in HBM when the timed region starts
  e2e     same metric through the array-level C ABI from pinned HOST buffers (reads + unit arrays copied H2D every step, results D2H)
  roofline  dominant kernel (k_build) against the measured HBM peak; cpu_baseline  the reference CPU path on a bounded sample
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

UNIT_BP = 4_600_000          # configs[1]: E. coli-sized unit
SAMPLE_BP = 1_150_000        # CPU legs: 1/4 of a unit, same shape (~6 s per pass on one host core of the GPU box)
SHAPE = dict(coverage=50, readlen=100, insert_mean=500, insert_sd=50, kmer=5, cov=20, contig_len=10000, contig_gap=1000, snp=0.01)
SEED = 20260925 + 2
METRIC = "graph-build+extend Mbp/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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


def make_samples(base, n):
    from oracle import harness
    from tools import synth
    harness.build_tools(with_emul=False)
    dirs = []
    for i in range(n):
        d = os.path.join(base, f"sample{i}")
        synth.synth(d, genome_bp=SAMPLE_BP, seed=SEED + 100 + i, **SHAPE)
        harness.prepare_tmp(d)
        dirs.append(d)
    return dirs


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    cores = min(n, os.cpu_count() or 1)
    base = tempfile.mkdtemp(prefix="ag_bench_ref_")
    try:
        dirs = make_samples(base, n)
        kind = "port"
        for _ in range(args.warmup):
            cpu_pass(dirs)
        t0 = time.perf_counter()
        worst = 0.0
        for _ in range(args.steps):
            s, kind = cpu_pass(dirs)
            worst += s
        wall = time.perf_counter() - t0
        value = n * SAMPLE_BP / 1e6 * args.steps / worst
        line = {
            "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "Mbp/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * worst / args.steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(n, sample=True),
            "cpu_baseline": {"value": round(value, 5), "unit": "Mbp/s", "cores": cores, "kind": kind,
                             "sample": f"{n} unit(s) of {SAMPLE_BP} bp (1/4 of the {UNIT_BP} bp unit, same shape), one process per unit; hot path = span "
                                       f"of the reference's progress lines (loadGenome..scaffoldContigs), text parsing included as in the reference"},
            "e2e": {"value": round(value, 5), "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(wall, 2),
        }
        emit(line)
    finally:
        shutil.rmtree(base, ignore_errors=True)


def workload_config(n, sample=False):
    bp = SAMPLE_BP if sample else UNIT_BP
    return {"workload": f"BASELINE configs[1] shape: {bp / 1e6:g} Mbp unit x {n} (one per GPU), 50x synthetic 2x100 bp PE (insert 500+-50), k=5, "
                        f"coverage=20, 10 kbp contig tiles, SNP 1%", "unit_bp": bp, "units": n, "pairs_per_unit": int(bp * 50 / 200), "readlen": 100, "k": 5,
            "coverage_threshold": 20, "insert_variation": 50, "parallelism": f"units{n} (1 unit/GPU, no data-path collective; one NCCL broadcast of the packed reads)",
            "l2": "inputs+tables per step (~1 GB) exceed the 126 MB L2; no explicit flush"}


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
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import aligngraph_b200 as ag
    from aligngraph_b200 import build as _build
    from tools import synth
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        _build.build()  # no-op when the in-tree library is newer than its sources

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE {world} != --gpus {args.gpus}")
    n = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # one process per GPU: give every rank its own block of host cores (post passes and staging are memory-bound host work;
        # unbound ranks pile up on one NUMA node)
        try:
            allowed = sorted(os.sched_getaffinity(0))
            mine = None
            try:   # cores local to this GPU (NVML), shared evenly by the ranks whose GPUs sit on the same NUMA node
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
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic job: rank 0 writes the n-unit data set once; every rank parses only its own unit --------------------------
    box = [None]
    if rank == 0:
        base = tempfile.mkdtemp(prefix="ag_bench_")
        t0 = time.perf_counter()
        synth.synth(base, genome_bp=UNIT_BP * n, chroms=n, seed=SEED, **SHAPE)
        log(f"[bench] synthetic data ({n} unit(s)) in {time.perf_counter() - t0:.1f} s -> {base}")
        box[0] = base
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    base = box[0]
    tmp = os.path.join(base, "tmp")
    ctx = ag.Context(k=SHAPE["kmer"], insert_variation=50, coverage=SHAPE["cov"], device=local)
    if rank == 0:
        ctx.formalize_inputs(os.path.join(base, "contigs.fa"), os.path.join(base, "genome.fa"), tmp, 1)
    barrier()

    # ---- reads: parsed once on rank 0, one NCCL broadcast of the packed buffer -------------------------------
    bcast = None
    t_reads = time.perf_counter()
    if world == 1:
        ctx.load_reads_fasta(os.path.join(tmp, "_reads.fa"))
    else:
        import numpy as np
        meta = [None]
        if rank == 0:
            ctx.load_reads_fasta(os.path.join(tmp, "_reads.fa"))
            b, m, l, npairs, s2, sm = ctx.get_reads()
            meta[0] = (npairs, s2, sm)
        dist.broadcast_object_list(meta, src=0)
        npairs, s2, sm = meta[0]
        nbytes_each = [2 * npairs * s2 * 4, 2 * npairs * sm * 4, npairs * 2]  # byte buffers: NCCL has no 16-bit integer type
        padded = [(cnt + world * 16 - 1) // (world * 16) * (world * 16) for cnt in nbytes_each]   # equal, 16-byte aligned slices for the all-gather
        dev = [torch.zeros(cnt, dtype=torch.uint8, device="cuda") for cnt in padded]
        if rank == 0:
            host = [np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(cnt,)) for p, cnt in zip((b, m, l), nbytes_each)]
            for t, h in zip(dev, host):
                t[:h.shape[0]].copy_(torch.from_numpy(h))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in dev:
            dist.broadcast(t, src=0)
        e1.record(); torch.cuda.synchronize()
        nbytes = sum(t.numel() * t.element_size() for t in dev)
        ms = e0.elapsed_time(e1)
        bcast = {"bytes": nbytes, "ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}
        ctx.set_reads(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), npairs, s2, sm, on_device=True)
        ctx._keep.append(dev)
        # end-to-end leg at N > 1: every rank owns 1/N of the packed read buffer in pinned host memory (in a deployment: the slice of
        # tmp/_reads.fa it parsed); per step it copies its slice H2D and one NCCL all-gather over NVLink rebuilds the full buffer on every
        # GPU — the bandwidth-optimal form of the read broadcast, with constant host->device bytes per rank
        slices = []
        for t in dev:
            n_sl = t.numel() // world
            view = t[rank * n_sl:(rank + 1) * n_sl]
            slices.append((view, view.cpu().pin_memory()))
    t_reads = time.perf_counter() - t_reads

    # ---- this rank's unit ----------------------------------------------------------------------------------------------------------
    unit = rank
    t0 = time.perf_counter()
    ctx.prepare_unit(tmp, unit)
    t_parse_unit = time.perf_counter() - t0
    ctx.pin_staged()

    def step():
        ctx.build()
        ctx.extend()

    for _ in range(max(args.warmup, 3)):
        step()
    check_pre = ctx.text(1)
    # ---- timed: inputs resident in HBM ------------------------------------------------------------------------------------------
    ctx.reset_stats()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    st = ctx.stats()
    barrier()
    # ---- timed: end to end through the array-level C ABI from pinned host buffers -------------------------------------------------
    ctx.reset_stats()
    barrier()
    ctx.timer_start()
    reads_h2d = 0
    for _ in range(args.steps):
        if world == 1:
            ctx.reupload_reads()
        else:
            for t, (view, host_slice) in zip(dev, slices):
                view.copy_(host_slice, non_blocking=True)
                dist.all_gather_into_tensor(t, view)
                reads_h2d += host_slice.numel()
            torch.cuda.synchronize()   # the library launches on its own stream
        ctx.invalidate_device_inputs()
        step()
    ms_e2e = ctx.timer_stop()
    st_e2e = ctx.stats()
    clock_info = clocks.stop() if clocks else None   # sampled every 100 ms across both timed regions
    barrier()
    assert ctx.text(1) == check_pre and len(check_pre) > 0
    # ---- file level (text in, text out), for reference: one pass --------------------------------------------------------------------
    t0 = time.perf_counter()
    ctx.prepare_unit(tmp, unit); ctx.build(); ctx.extend(); ctx.write_unit(tmp, unit)
    t_file = time.perf_counter() - t0

    t = torch.tensor([ms, ms_e2e, t_file, t_parse_unit], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, t_file, t_parse_unit = t.tolist()

    if rank == 0:
        K = args.steps
        mbp = n * UNIT_BP / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        # dominant kernel: k_build (node sweep + common-case edges).  Algorithmic bytes of one launch: per tile key the
        # alignment index, its 32-byte prepared record and the left mate's packed bases + non-ACGT plane; per position the contiMer
        # summary, reference base and CSR entry; per node the final records written (16 + 16 + 8 + 4).
        L = SHAPE["readlen"]
        alg = st["n_keys"] * (4 + 32 + (L + 3) // 4 + (L + 7) // 8) + UNIT_BP * (8 + 1 + 4) + st["n_nodes"] * 44
        nodes_ms = st["ms_nodes"] / K
        achieved = alg / nodes_ms / 1e6
        pairs = UNIT_BP * 50 // 200
        k1_equiv = pairs * (2 * ((L + 3) // 4) + 32 + 16 * (L - SHAPE["kmer"])) / ((st["ms_prep"] + st["ms_sort"] + st["ms_nodes"]) / K) / 1e6
        traffic = None
        try:   # DRAM bytes of one k_build launch from the committed `ncu --set full` capture of this same workload (profiles/README.md)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["traffic"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": round(mbp * K / (ms / 1000), 3), "unit": "Mbp/s", "n_gpus": n, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(n),
            "e2e": {"value": round(mbp * K / (ms_e2e / 1000), 3), "unit": "Mbp/s", "h2d_bytes_per_step": int((st_e2e["h2d_bytes"] + reads_h2d) / K),
                    "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / K), "ms_per_step": round(ms_e2e / K, 3)},
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": "k_build", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "alg_bytes_per_launch": int(alg), "ms_per_launch": round(nodes_ms, 4),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         "survey_k1_equiv_gbs": round(k1_equiv, 1)},
            "device_ms_per_step": {k[3:]: round(st[k] / K, 4) for k in st if k.startswith("ms_")},
            "counts_per_unit": {"alignments": int(st["n_aln"] / K), "nodes": int(st["n_nodes"]), "walks": int(st["n_walks"] / K),
                                "emitted_contigs": int(st["n_emitted"] / K), "tile_keys": int(st["n_keys"])},
            "host_s_per_step": {"device_section": round(st["s_device_section"] / K, 4), "post_passes": round(st["s_post"] / K, 4)},
            "file_level": {"value": round(mbp / t_file, 3), "unit": "Mbp/s", "s_per_unit": round(t_file, 3),
                           "note": "tmp/ text files in -> tmp/ FASTA out through ag_run_unit_files steps (SAM/PSL parse + device + write), reads parsed once: "
                                   f"{t_reads:.2f} s extra", "parse_unit_s": round(t_parse_unit, 3)},
            "reads_broadcast": bcast,
            "walk_fallback": int(st["walk_fallback"]),
        }
        if not args.no_cpu and n == 1:
            try:
                sdir = tempfile.mkdtemp(prefix="ag_bench_cpu_")
                dirs = make_samples(sdir, 1)
                s, kind = cpu_pass(dirs)
                line["cpu_baseline"] = {"value": round(SAMPLE_BP / 1e6 / s, 5), "unit": "Mbp/s", "cores": 1, "kind": kind,
                                        "sample": f"one {SAMPLE_BP} bp unit (1/4 of the workload unit, same shape), hot path of the reference's CPU "
                                                  f"implementation timed once: {s:.2f} s"}
                shutil.rmtree(sdir, ignore_errors=True)
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "Mbp/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    barrier()
    ctx.close()
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()

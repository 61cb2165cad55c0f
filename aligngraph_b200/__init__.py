"""aligngraph_b200 — B200-native drop-in for AlignGraph's per-chromosome graph-build + contig-extension path.

Python is only the test / benchmark harness: this module is a thin ctypes binding of the C ABI declared in
``include/aligngraph_b200.h`` (the product is the shared library + the ``AlignGraph`` CLI).  There is no CPU fallback: if the
CUDA library is missing or no GPU is present, construction of a :class:`Context` raises.

Method names mirror the reference's loop body (AlignGraph.cpp:4768-4776): ``run_unit`` = loadGenome + loadContigAlignment +
loadReadAlignment + extendContigs + scaffoldContigs for one unit, on the reference's own tmp/ files.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AG_LIB_PATH") or os.path.join(_HERE, "libaligngraph_b200.so")  # AG_LIB_PATH: tuning variants only
HEADER_PATH = os.path.join(_HERE, "..", "include", "aligngraph_b200.h")


class AlignGraphError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("k", C.c_int), ("insert_variation", C.c_int), ("coverage", C.c_int), ("device", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_prep", "ms_sort", "ms_nodes", "ms_finalize", "ms_edges", "ms_components",
                                         "ms_chains", "ms_walk", "ms_materialize", "ms_d2h")] + \
               [(n, C.c_double) for n in ("s_parse", "s_device_section", "s_post")] + \
               [(n, C.c_uint64) for n in ("n_aln", "n_nodes", "n_walks", "n_emitted", "n_keys", "n_tiles", "kernel_launches",
                                          "h2d_bytes", "d2h_bytes")] + \
               [("walk_fallback", C.c_int), ("ms_ingest_reads", C.c_float), ("ms_ingest_sam", C.c_float)] + \
               [(n, C.c_uint64) for n in ("sam_device", "sam_host", "reads_device", "reads_host", "regrows", "reads_windowed")] + \
               [("ms_stage", C.c_float), ("ms_build_kernel", C.c_float), ("ms_select", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class UnitView(C.Structure):
    _fields_ = [("ref", C.c_void_p), ("n_ref", C.c_uint32), ("n_tail", C.c_uint32), ("cm_start", C.c_void_p), ("cm", C.c_void_p),
                ("n_cm", C.c_uint32), ("chain_pos", C.c_void_p), ("chain_base", C.c_void_p), ("aln", C.c_void_p), ("n_aln", C.c_uint64),
                ("ext", C.c_void_p), ("n_ext", C.c_uint64), ("threads", C.c_void_p), ("n_threads", C.c_uint32)]


_lib = None


def load_library(path=None):
    """Load the CUDA shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise AlignGraphError(f"{path} not found: build it with `python -m aligngraph_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, cp, i32, u32, u64 = C.c_void_p, C.c_char_p, C.c_int, C.c_uint32, C.c_uint64
    sig = {
        "ag_create": (i32, [C.POINTER(Params), C.POINTER(vp)]),
        "ag_destroy": (None, [vp]),
        "ag_last_error": (cp, [vp]),
        "ag_create_error": (cp, []),
        "ag_set_reads": (i32, [vp, vp, vp, vp, u64, u32, u32]),
        "ag_set_reads_device": (i32, [vp, vp, vp, vp, u64, u32, u32]),
        "ag_set_read_exceptions": (i32, [vp, vp, vp, u64]),
        "ag_load_reads_fasta": (i32, [vp, cp]),
        "ag_load_reads_for_units": (i32, [vp, cp, cp, C.POINTER(i32), i32]),
        "ag_get_reads": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]),
        "ag_broadcast_reads": (i32, [C.POINTER(vp), i32, vp]),
        "ag_begin_unit": (i32, [vp, i32, vp, u32]),
        "ag_set_contimers": (i32, [vp, vp, vp, u32, vp, vp, vp, u32]),
        "ag_set_contig_threads": (i32, [vp, vp, u32, vp, vp, u32, vp, u32]),
        "ag_add_alignments": (i32, [vp, vp, u64, vp, u64]),
        "ag_build": (i32, [vp]),
        "ag_extend": (i32, [vp]),
        "ag_process": (i32, [vp]),
        "ag_get_text": (i32, [vp, i32, C.POINTER(vp), C.POINTER(u64)]),
        "ag_prepare_unit_files": (i32, [vp, cp, i32]),
        "ag_write_unit_files": (i32, [vp, cp, i32]),
        "ag_run_unit_files": (i32, [vp, cp, i32]),
        "ag_run_units_files": (i32, [C.POINTER(vp), i32, cp, i32, i32, i32, vp, vp]),
        "ag_run_job_files": (i32, [C.POINTER(vp), i32, cp, cp, C.POINTER(i32), i32, i32, vp, vp]),
        "ag_get_unit": (i32, [vp, C.POINTER(UnitView)]),
        "ag_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "ag_reset_stats": (i32, [vp]),
        "ag_keep_node_counts": (i32, [vp, i32]),
        "ag_dump_nodes_text": (i32, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "ag_cuda_stream": (vp, [vp]),
        "ag_invalidate_device_inputs": (i32, [vp]),
        "ag_reupload_reads": (i32, [vp]),
        "ag_pin_staged": (i32, [vp]),
        "ag_formalize_inputs": (i32, [vp, cp, cp, cp, i32, C.POINTER(i32)]),
        "ag_set_option": (i32, [vp, cp, C.c_long]),
        "ag_remove_misassembly_file": (i32, [vp, cp, cp, i32, cp, vp, vp]),
        "ag_containment_search_files": (i32, [vp, cp, cp, cp]),
        "ag_timer_start": (i32, [vp]),
        "ag_timer_stop": (i32, [vp, C.POINTER(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._ag_signatures = sig
    if path == LIB_PATH:
        _lib = lib
    return lib


class Context:
    """One GPU context (``ag_ctx``).  ``k`` / ``insert_variation`` / ``coverage`` are --kMer / --insertVariation / --coverage."""

    def __init__(self, k=5, insert_variation=50, coverage=20, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        p = Params(k, insert_variation, coverage, device)
        rc = self._lib.ag_create(C.byref(p), C.byref(self._h))
        if rc != 0:
            raise AlignGraphError("ag_create: " + self._lib.ag_create_error().decode())
        self._keep = []

    def close(self):
        if self._h:
            self._lib.ag_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise AlignGraphError(f"{what}: {self._lib.ag_last_error(self._h).decode()}")

    # ---- reads ---------------------------------------------------------------------------------------------------------
    def load_reads_fasta(self, path):
        self._ck(self._lib.ag_load_reads_fasta(self._h, os.fsencode(path)), "ag_load_reads_fasta")

    def load_reads_for_units(self, reads_fa, tmp_dir, units):
        ul = (C.c_int * len(units))(*units)
        self._ck(self._lib.ag_load_reads_for_units(self._h, os.fsencode(reads_fa), os.fsencode(tmp_dir), ul, len(units)), "ag_load_reads_for_units")

    def get_reads(self):
        """(bases_ptr, nmask_ptr, len_ptr, n_pairs, stride2, stridem) of the packed host copy."""
        b, m, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, s2, sm = C.c_uint64(), C.c_uint32(), C.c_uint32()
        self._ck(self._lib.ag_get_reads(self._h, C.byref(b), C.byref(m), C.byref(l), C.byref(n), C.byref(s2), C.byref(sm)), "ag_get_reads")
        return b.value, m.value, l.value, n.value, s2.value, sm.value

    def set_reads(self, bases_ptr, nmask_ptr, len_ptr, n_pairs, stride2, stridem, on_device=False):
        fn = self._lib.ag_set_reads_device if on_device else self._lib.ag_set_reads
        self._ck(fn(self._h, bases_ptr, nmask_ptr, len_ptr, n_pairs, stride2, stridem), "ag_set_reads")

    # ---- file level (the reference's loop body) ---------------------------------------------------------------------------
    def run_unit(self, tmp_dir, unit):
        self._ck(self._lib.ag_run_unit_files(self._h, os.fsencode(tmp_dir), unit), "ag_run_unit_files")

    def run_units(self, tmp_dir, first, n, prefetch=4):
        """Units [first, first+n) with the host parsing pipelined ahead of the GPU (ag_run_units_files, this context only)."""
        arr = (C.c_void_p * 1)(self._h)
        rc = self._lib.ag_run_units_files(arr, 1, os.fsencode(tmp_dir), first, n, prefetch, None, None)
        self._ck(rc, "ag_run_units_files")

    def run_job(self, tmp_dir, units, reads_fa=None, prefetch=2):
        """The whole hot loop for `units` (ag_run_job_files, this context only); reads_fa: (re)load the read set as part of the job."""
        arr = (C.c_void_p * 1)(self._h)
        ul = (C.c_int * len(units))(*units)
        rc = self._lib.ag_run_job_files(arr, 1, os.fsencode(tmp_dir), os.fsencode(reads_fa) if reads_fa else None, ul, len(units), prefetch, None, None)
        self._ck(rc, "ag_run_job_files")

    @staticmethod
    def run_job_on(ctxs, tmp_dir, units, reads_fa=None, prefetch=2):
        """ag_run_job_files over several contexts (one worker per context; contexts may share a GPU: a second context on the same device
        overlaps its unit's text staging and host post passes with the first one's kernels).  The reads are loaded into ctxs[0] and copied."""
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        ul = (C.c_int * len(units))(*units)
        rc = ctxs[0]._lib.ag_run_job_files(arr, len(ctxs), os.fsencode(tmp_dir), os.fsencode(reads_fa) if reads_fa else None, ul, len(units), prefetch, None, None)
        if rc:
            msgs = [c._lib.ag_last_error(c._h).decode() for c in ctxs]
            raise AlignGraphError("ag_run_job_files: " + next((m for m in msgs if m), "failed"))

    def prepare_unit(self, tmp_dir, unit):
        self._ck(self._lib.ag_prepare_unit_files(self._h, os.fsencode(tmp_dir), unit), "ag_prepare_unit_files")

    def write_unit(self, tmp_dir, unit):
        self._ck(self._lib.ag_write_unit_files(self._h, os.fsencode(tmp_dir), unit), "ag_write_unit_files")

    # ---- array level ----------------------------------------------------------------------------------------------------------
    def get_unit(self):
        v = UnitView()
        self._ck(self._lib.ag_get_unit(self._h, C.byref(v)), "ag_get_unit")
        return v

    def begin_unit(self, unit, ref_ptr, n_ref):
        self._ck(self._lib.ag_begin_unit(self._h, unit, ref_ptr, n_ref), "ag_begin_unit")

    def set_contig_threads(self, threads, n_threads, chain_pos, chain_base, n_cm, tail_ptr, n_tail):
        self._ck(self._lib.ag_set_contig_threads(self._h, threads, n_threads, chain_pos, chain_base, n_cm, tail_ptr, n_tail), "ag_set_contig_threads")

    def set_contimers(self, cm_start, cm, n_cm, chain_pos, chain_base, tail_ptr, n_tail):
        self._ck(self._lib.ag_set_contimers(self._h, cm_start, cm, n_cm, chain_pos, chain_base, tail_ptr, n_tail), "ag_set_contimers")

    def add_alignments(self, aln, n, ext, n_ext):
        self._ck(self._lib.ag_add_alignments(self._h, aln, n, ext, n_ext), "ag_add_alignments")

    def build(self):
        self._ck(self._lib.ag_build(self._h), "ag_build")

    def extend(self):
        self._ck(self._lib.ag_extend(self._h), "ag_extend")

    def process(self):
        """build + extend as one step with a single host synchronisation (ag_process)."""
        self._ck(self._lib.ag_process(self._h), "ag_process")

    def text(self, which):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.ag_get_text(self._h, which, C.byref(p), C.byref(n)), "ag_get_text")
        return C.string_at(p.value, n.value) if n.value else b""

    # ---- introspection ----------------------------------------------------------------------------------------------------------
    def stats(self):
        s = Stats()
        self._ck(self._lib.ag_get_stats(self._h, C.byref(s)), "ag_get_stats")
        return s.as_dict()

    def reset_stats(self):
        self._ck(self._lib.ag_reset_stats(self._h), "ag_reset_stats")

    def keep_node_counts(self, on=True):
        """Keep coverage / base counters per node after the build (needed by dump_nodes_text; tests only)."""
        self._ck(self._lib.ag_keep_node_counts(self._h, 1 if on else 0), "ag_keep_node_counts")

    def dump_nodes_text(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.ag_dump_nodes_text(self._h, C.byref(p), C.byref(n)), "ag_dump_nodes_text")
        return C.string_at(p.value, n.value) if n.value else b""

    def invalidate_device_inputs(self):
        self._ck(self._lib.ag_invalidate_device_inputs(self._h), "ag_invalidate_device_inputs")

    def reupload_reads(self):
        self._ck(self._lib.ag_reupload_reads(self._h), "ag_reupload_reads")

    def pin_staged(self):
        self._ck(self._lib.ag_pin_staged(self._h), "ag_pin_staged")

    def formalize_inputs(self, contig_fa, genome_fa, tmp_dir, part=1):
        n = C.c_int()
        self._ck(self._lib.ag_formalize_inputs(self._h, os.fsencode(contig_fa), os.fsencode(genome_fa), os.fsencode(tmp_dir), part, C.byref(n)),
                 "ag_formalize_inputs")
        return n.value

    def set_option(self, name, value):
        self._ck(self._lib.ag_set_option(self._h, name.encode(), int(value)), "ag_set_option")

    def timer_start(self):
        self._ck(self._lib.ag_timer_start(self._h), "ag_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self._lib.ag_timer_stop(self._h, C.byref(ms)), "ag_timer_stop")
        return ms.value

    def cuda_stream(self):
        return self._lib.ag_cuda_stream(self._h)

"""Build tuning variants of the library (compile-time knobs of the tile sweeps) into aligngraph_b200/_variants/lib_<name>.so;
tools/variants.sh then runs bench.py against each of them on the GPU box (AG_LIB_PATH selects the library)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aligngraph_b200 import build as b  # noqa: E402

VARIANTS = {
    "tile7": ["-DAG_TWARPS=7"],                                  # 7-warp tiles (217 positions): 56 registers at 5 CTAs per SM
    "tma_minb5": ["-DAG_TMA_MINB=5"],                            # TMA-staged node sweep at 48 registers / 5 CTAs per SM
    "tma_minb4": ["-DAG_TMA_MINB=4"],                            # ... at 64 registers / 4 CTAs per SM
    "code4": ["-DAG_CODE4=1"],                                   # left mates staged as oriented 4-bit codes (CPU-verified coder, not yet timed)
    "code4_minb5": ["-DAG_CODE4=1", "-DAG_NODES_MINB=5"],
    "minb5": ["-DAG_NODES_MINB=5"],
    "minb6_chunk64": ["-DAG_NODES_MINB=6", "-DAG_NCHUNK_NODES=64"],
}


def main(names):
    out = os.path.join(b.HERE, "_variants")
    os.makedirs(out, exist_ok=True)
    for name in names or VARIANTS:
        lib = os.path.join(out, f"lib_{name}.so")
        cmd = ["nvcc", *b.ARCH, *b.COMMON, *VARIANTS[name], "-shared", "-o", lib,
               os.path.join(b.CSRC, "ag_device.cu"), os.path.join(b.CSRC, "ag_host.cpp"), os.path.join(b.CSRC, "ag_capi.cpp")]
        subprocess.run(cmd, check=True)
        print(lib)


if __name__ == "__main__":
    main(sys.argv[1:])

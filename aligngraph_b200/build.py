"""Build the sm_100a shared library (and the drop-in CLI) in-tree with nvcc.

    python -m aligngraph_b200.build          # -> aligngraph_b200/libaligngraph_b200.so, aligngraph_b200/bin/AlignGraph

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaligngraph_b200.so")
CLI = os.path.join(HERE, "bin", "AlignGraph")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-pthread"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(HERE, "..", "include", "aligngraph_b200.h"))
    return files


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    srcs = _sources()
    if force or _newer(LIB, srcs):
        cmd = [nvcc, *ARCH, *COMMON, "-shared", "-o", LIB,
               os.path.join(CSRC, "ag_device.cu"), os.path.join(CSRC, "ag_host.cpp"), os.path.join(CSRC, "ag_capi.cpp")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    main_cpp = os.path.join(CSRC, "ag_main.cpp")
    if os.path.exists(main_cpp) and (force or _newer(CLI, srcs + [LIB])):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        subprocess.run([nvcc, *ARCH, *COMMON, "-o", CLI, main_cpp, "-L" + HERE, "-laligngraph_b200",
                        "-Xlinker", "-rpath=$ORIGIN/.."], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

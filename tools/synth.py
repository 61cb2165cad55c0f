"""Build and run the deterministic synthetic workload generator (tools/agsynth.cpp).  Data generation only — no oracle code."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_bin", "agsynth")


def build():
    src = os.path.join(HERE, "agsynth.cpp")
    if not os.path.exists(BIN) or os.path.getmtime(src) > os.path.getmtime(BIN):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-o", BIN, src], check=True)
    return BIN


def synth(out_dir, **params):
    """Generate a work directory; params map to agsynth options (underscores -> dashes).  Returns the meta dict."""
    cmd = [build(), "--out", out_dir]
    for k, v in params.items():
        cmd += ["--" + k.replace("_", "-"), str(v)]
    subprocess.run(cmd, check=True)
    meta = {}
    with open(os.path.join(out_dir, "synth_meta.txt")) as f:
        for line in f:
            k, v = line.split()
            meta[k] = int(v)
    return meta

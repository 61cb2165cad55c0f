mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r03_bench_n2.json 2> gpurun_out/r03_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['device_ms_per_step'])"
tail -3 gpurun_out/r03_bench_n2.err
# product path on two GPUs: the CLI farms units over AG_DEVICES, one NCCL broadcast of the packed reads
python - <<'PY'
import os, sys, subprocess, tempfile, shutil
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from oracle import harness
import cases
harness.build_tools(with_ref=False, with_emul=False)
w = tempfile.mkdtemp(prefix='ag_two_')
harness.synth(w, **cases.GOLDEN['two_chr'])
env = harness.stub_env(); env['AG_DEVICES'] = '0,1'; env['AG_STATS'] = '1'
r = subprocess.run([os.path.join(os.getcwd(), 'aligngraph_b200', 'bin', 'AlignGraph'), '--resume'], cwd=w, env=env, capture_output=True, text=True, timeout=600)
print('cli rc', r.returncode); print(r.stderr[-900:])
g = os.path.join('tests', 'golden', 'two_chr')
for f in ('extendedContigs.fa', 'remainingContigs.fa'):
    print(f, open(os.path.join(w, f), 'rb').read() == open(os.path.join(g, f), 'rb').read())
shutil.rmtree(w)
PY
timeout 600 python -m pytest tests/test_containment.py tests/test_gpu_parity.py -m gpu -x -q -k "not full_size_unit" 2>&1 | grep -v "^$" | tail -4

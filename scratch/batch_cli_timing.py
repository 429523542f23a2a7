"""Throughput of the C++ all-vs-all driver (oracle/_ref/strict/minrms_batch_cuda, one process per
pair) next to the Python driver (minrms_b200.batch, host threads sharing a process) on the same
synthetic chains (config-5 shape: lengths 150..400, seeds 5000 + i), one GPU.
python scratch/batch_cli_timing.py [chains=10] [K ...]"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minrms_b200.synth import random_chain, write_ca_pdb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ks = [int(x) for x in sys.argv[2:]] or [1, 4]
d = tempfile.mkdtemp()
names = []
for i in range(n):
    rng = np.random.default_rng(5000 + i)
    L = int(rng.integers(150, 401))
    write_ca_pdb(os.path.join(d, "c%03d.pdb" % i), random_chain(rng, L))
    names.append("c%03d.pdb" % i)
open(os.path.join(d, "list.txt"), "w").write("\n".join(names) + "\n")
pairs = n * (n - 1) // 2
exe = os.path.join(ROOT, "oracle", "_ref", "strict", "minrms_batch_cuda")
for k in ks:
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-fm", "4", "-HS", "-ir", "-r", "-batch", "list.txt", "-batch-out", "out%d" % k,
                        "-batch-workers-per-gpu", str(k)], cwd=d, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    print("C++ driver, %d chains, %d pairs, %d slot(s): %.1f s, %.2f pairs/s, rc %d | %s" % (
        n, pairs, k, dt, pairs / dt, r.returncode, r.stdout.strip().splitlines()[-1]))
if os.environ.get("SKIP_PY"):
    sys.exit(0)
t0 = time.perf_counter()
r = subprocess.run([sys.executable, "-m", "minrms_b200.batch", d, "--out", os.path.join(d, "py.npz"), "--msf-dir",
                    os.path.join(d, "pyout")], cwd=ROOT, capture_output=True, text=True)
dt = time.perf_counter() - t0
print("Python driver (4 threads, GPU orientation generation): %.1f s, %.2f pairs/s, rc %d" % (dt, pairs / dt, r.returncode))

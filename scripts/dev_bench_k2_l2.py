"""Config-2 device loop (persistent euclidean k-centers kernel) as a function of the part of the
shard kept L2-resident across iterations (EB_K2_L2_MB).  One process per setting (the library
reads the switch once)."""
import os
import subprocess
import sys

CHILD = r'''
import torch
from enspara_b200 import synth
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster.kcenters import _SingleComm
import os
N = int(os.environ.get("EB_N", "1000000"))
X = synth.device_features(N, 64, seed=0)
best = 1e9
for rep in range(3):
    eng = KCentersEngine(X, "euclidean", _SingleComm())
    eng.run(20, 0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); eng.run(520, 0.0); e1.record(); torch.cuda.synchronize()
    best = min(best, 1e3 * e0.elapsed_time(e1) / 500)
st = eng.read_state()
import hashlib
h = hashlib.sha1(eng.assign.cpu().numpy().tobytes() + eng.dist.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"{best:.2f} us/iter  barrier {1e-3*st.wait_ns/519:.2f} us  sha {h}")
'''
for mb in sys.argv[1:] or ["0", "32", "48", "64", "80", "96", "128"]:
    env = dict(os.environ, EB_K2_L2_MB=mb)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"N={os.environ.get('EB_N', '1000000')} EB_K2_L2_MB={mb}: {out.stdout.strip() or out.stderr[-400:]}", flush=True)

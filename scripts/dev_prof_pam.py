"""Developer probe: where a PAM proposal's time goes (phase timings with device syncs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster import util, kcenters as kc
from enspara_b200.cluster._pam import PamEngine
from enspara_b200.cluster.kcenters import _SingleComm

n, A, k, nprop = 1_000_000, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 100
data = synth.device_trajectory(n, A, seed=0)
res, eng = kc.kcenters(data, "rmsd", n_clusters=k, _return_engine=True)
for use_tc in (True, False):
    pam = PamEngine(data, util.RMSD, _SingleComm(), eng.dist, eng.assign,
                    [int(c) for c in res.center_indices])
    pam.use_tc = use_tc and pam.use_tc
    pam.sweep(random_state=0, max_proposals=5)
    pam.profile = {}
    acc = pam.sweep(random_state=1, max_proposals=nprop)
    out = {kk: (1e3 * v / nprop if kk != "n_ambig" else v / nprop) for kk, v in pam.profile.items()}
    print("use_tc=%s accepted=%d per-proposal ms:" % (pam.use_tc, acc), json.dumps(out), flush=True)

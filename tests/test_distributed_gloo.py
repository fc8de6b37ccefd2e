"""CPU: world_size-2 gloo run of the shard-level host logic (enspara_b200.mpi, ShardInfo,
index conversions, candidate-record exchange)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29577",
           os.path.join(HERE, "gloo_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "GLOO_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]

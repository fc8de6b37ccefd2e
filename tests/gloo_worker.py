"""world_size-2 gloo worker (CPU): the shard-level host logic of the N>1 path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Traj:
    def __init__(self, xyz, topology=None):
        self.xyz = np.asarray(xyz, dtype=np.float32)
        if self.xyz.ndim == 2:
            self.xyz = self.xyz[None]
        self.top = topology

    def __len__(self):
        return len(self.xyz)

    def __getitem__(self, i):
        return Traj(self.xyz[i] if not isinstance(i, (int, np.integer)) else self.xyz[i][None],
                    self.top)


def main():
    import torch
    import torch.distributed as dist
    from enspara_b200 import mpi
    from enspara_b200.cluster._engine import ShardInfo
    from enspara_b200.cluster import kmedoids
    mpi.init_from_env("gloo")
    rank, size = mpi.rank(), mpi.size()
    assert size == 2 and rank in (0, 1)
    checks = []

    # striped max / mean (mpi/ops.py:128-166): data[rank::size] as in the reference's tests
    full = np.arange(1, 12, dtype=np.float64) ** 1.5
    mine = full[rank::size]
    checks.append(abs(mpi.ops.striped_array_max(mine) - full.max()) < 1e-12)
    checks.append(abs(mpi.ops.striped_array_mean(mine) - full.mean()) < 1e-12)

    # distribute_frame (ops.py:169-212), arrays and trajectory-likes
    data = np.arange(20, dtype=np.float64).reshape(10, 2) + 100 * rank
    fr = mpi.ops.distribute_frame(data, world_index=3, owner_rank=1)
    checks.append(np.array_equal(fr, np.array([106.0, 107.0])))
    t = Traj(np.arange(5 * 4 * 3, dtype=np.float32).reshape(5, 4, 3) + 1000 * rank, "top")
    fr = mpi.ops.distribute_frame(t, world_index=2, owner_rank=0)
    checks.append(isinstance(fr, Traj) and fr.top == "top" and
                  np.array_equal(fr.xyz, (np.arange(60, dtype=np.float32).reshape(5, 4, 3))[2:3]))

    # randind (ops.py:215-272): rank 0 draws, striped-concat map; every draw must be a valid
    # (owner, local) and the empirical choice must match the reference's formula
    local = np.arange(7 if rank == 0 else 4)
    rs = np.random.RandomState(0)
    ref_rs = np.random.RandomState(0)
    for _ in range(20):
        owner, li = mpi.ops.randind(local, rs)
        g = ref_rs.randint(11)
        concat = np.concatenate([np.arange(11)[r::2] for r in range(2)])
        pos = int(np.where(concat == g)[0][0])
        want = (0, pos) if pos < 7 else (1, pos - 7)
        checks.append((owner, li) == want)

    # assembly helpers (ops.py:42-125, 14-39): trajectory i lives on rank i % size
    lengths = np.array([3, 5, 2, 4])
    glob = np.arange(lengths.sum()) + 1
    starts = np.concatenate([[0], np.cumsum(lengths)])
    mine = np.concatenate([glob[starts[t]:starts[t + 1]] for t in range(rank, 4, 2)])
    checks.append(np.array_equal(mpi.ops.assemble_striped_ragged_array(mine, lengths), glob))
    checks.append(np.array_equal(mpi.ops.assemble_striped_array(lengths[rank::2]), lengths))
    conv = mpi.ops.convert_local_indices([(0, 0), (0, 4), (1, 0), (1, 6)], lengths)
    checks.append([int(c) for c in conv] == [0, 9, 3, 11])
    checks.append(kmedoids.ctr_ids_mpi([0, 9, 3, 11], lengths) == [(0, 0), (0, 4), (1, 0), (1, 6)])
    checks.append(kmedoids.ctr_ids_mpi([(2, 1), (3, 0)], lengths) == [(0, 4), (1, 5)])

    # shard bookkeeping of the device loops: contiguous blocks by rank
    sh = ShardInfo(5 if rank == 0 else 8, mpi.comm)
    checks.append(sh.n_global == 13 and sh.offset == (0 if rank == 0 else 5))
    checks.append(sh.to_rank_local(4) == (0, 4) and sh.to_rank_local(5) == (1, 0)
                  and sh.to_rank_local(12) == (1, 7))

    # candidate-record exchange: all-gather of fixed-size byte records, then the same winner
    # on every rank (max distance, lowest global index)
    rec = torch.zeros(48, dtype=torch.uint8)
    hdr = np.zeros(1, dtype=[("d", "f8"), ("i", "i8"), ("t", "f8"), ("r", "i8")])
    hdr["d"], hdr["i"] = (2.5, 3) if rank == 0 else (2.5, 9)
    rec[:32] = torch.from_numpy(hdr.view(np.uint8).copy())
    allrec = torch.zeros(96, dtype=torch.uint8)
    mpi.comm.all_gather_into(allrec, rec)
    got = allrec.numpy().reshape(2, 48)[:, :32].copy().view(hdr.dtype).reshape(2)
    best = max(range(2), key=lambda r: (got[r]["d"], -got[r]["i"]))
    checks.append(best == 0 and int(got[best]["i"]) == 3)

    ok = all(bool(c) for c in checks)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GLOO_OK" if int(flag) == 1 else "GLOO_FAIL %s" % checks, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()

"""Host driver of the fused k-centers device loop.

One iteration of the reference (/root/reference/enspara/cluster/kcenters.py:243-311 serial,
:314-378 per shard) is ONE kernel launch here (csrc/eb_rmsd_kcenters.cu, csrc/eb_feat.cu), plus
ONE all-gather of candidate records when frames are sharded over ranks.  The winner of the
cross-shard arg-max is chosen on the device by the next launch's prologue, so the host never
waits for it: launches are queued in batches and only a 64-byte state block is read back per
batch to learn whether the stop rule (kcenters.py:217) fired.
"""
import numpy as np
import torch

from .. import _lib
from ..device import DeviceFeatures, DeviceTrajectory, ptr, stream_ptr

_INT32_MAX = 2 ** 31 - 1


class ShardInfo:
    """Where this rank's frames sit in the global order: contiguous blocks by rank, so that
    'lowest rank, then lowest local index' (kcenters.py:337-338) == lowest global index.

    ``extra`` is an int this rank contributes alongside its length (the engine's "peer-memory
    exchange is set up" bit), so ONE small tensor all-gather carries everything the ranks have
    to agree on -- no pickled host collectives."""

    def __init__(self, n_local, comm, extra=0, device=None):
        self.comm = comm
        self.size = comm.size
        self.rank = comm.rank
        if self.size == 1:
            lens, extras = [int(n_local)], [int(extra)]
        else:
            both = gather_int_pairs(comm, int(n_local), int(extra), device)
            lens, extras = both[:, 0].tolist(), both[:, 1].tolist()
        self.extras = [int(e) for e in extras]
        self.lengths = np.asarray(lens, dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        self.offset = int(self.offsets[self.rank])
        self.n_global = int(self.offsets[-1])

    def to_rank_local(self, global_idx):
        r = int(np.searchsorted(self.offsets, global_idx, side="right") - 1)
        return r, int(global_idx - self.offsets[r])


def gather_int_pairs(comm, a, b, device=None):
    """(size, 2) int64 numpy array of every rank's (a, b): one tensor all-gather (NCCL on the
    device, gloo on the host) instead of ``all_gather_object``'s pickle + two collectives."""
    import torch.distributed as dist
    on_gpu = dist.get_backend(getattr(comm, "group", None)) == "nccl"
    if on_gpu:
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    mine = torch.tensor([int(a), int(b)], dtype=torch.int64, device=dev)
    out = torch.empty(2 * comm.size, dtype=torch.int64, device=dev)
    comm.all_gather_into(out, mine)
    return out.cpu().numpy().reshape(comm.size, 2)


class _ExchangePool:
    """Symmetric exchange buffers of the fused peer-memory candidate exchange, kept per
    (process group, buffer size).  Allocation + rendezvous of torch symmetric memory costs
    0.1-0.3 s (CUDA VMM mapping into every peer), so an engine borrows a buffer for a run and
    hands it back when the run ended cleanly; the sequence numbers inside simply continue
    (a flag is compared for equality with the consumer's own publish counter, and every rank
    makes the same sequence of publishes).  A run that raised keeps its buffer, which is then
    never reused."""

    def __init__(self):
        self.free = {}

    def take(self, key):
        lst = self.free.get(key)
        return lst.pop() if lst else None

    def give(self, key, item):
        self.free.setdefault(key, []).append(item)


_exchange_pool = _ExchangePool()


class KCentersEngine:
    """Device state of one k-centers run over one shard."""

    #: launches queued between two reads of the state block when a distance cutoff is active
    POLL_EVERY = 32

    def __init__(self, data, metric_kind, comm, exact=True, triangle=False):
        self.data = data
        self.kind = metric_kind
        self.comm = comm
        self.exact = bool(exact)
        self.lib = _lib.load()
        self.n = len(data)
        self.is_rmsd = isinstance(data, DeviceTrajectory)
        dev = data.xyz.device if self.is_rmsd else data.X.device
        self.dev = dev
        # use_triangle_inequality (kcenters.py:287-296): RMSD only; needs the coordinates of the
        # centres chosen so far (center_store) and the new centre's distance to each (cc)
        self.triangle = bool(triangle) and self.is_rmsd and self.exact
        # sharded RMSD runs exchange the candidate records through peer memory inside the step
        # kernel (no collective launch); anything else uses one NCCL all-gather per step.
        # Every rank must take the same path: the "set up" bit travels with the shard lengths.
        p2p_ready = self._setup_p2p()
        self.shard = ShardInfo(self.n, comm, extra=int(p2p_ready), device=dev)
        self.p2p = bool(p2p_ready) and all(self.shard.extras)
        if p2p_ready and not self.p2p:
            self._release_p2p(reusable=True)
        if self.is_rmsd:
            self.rec_bytes = int(self.lib.eb_rmsd_record_bytes(data.n_atoms))
            self.dist = torch.full((self.n,), float("inf"), dtype=torch.float32, device=dev)
        else:
            self.rec_bytes = int(self.lib.eb_feat_record_bytes(data.n_features, data.dt))
            self.dist = torch.full((self.n,), float("inf"), dtype=torch.float64, device=dev)
        self.assign = torch.full((self.n,), -1, dtype=torch.int32, device=dev)
        self.state = torch.zeros(64, dtype=torch.uint8, device=dev)
        self.partials = torch.empty(int(self.lib.eb_kc_partials_bytes()), dtype=torch.uint8,
                                    device=dev)
        self.cand_out = torch.zeros(self.rec_bytes, dtype=torch.uint8, device=dev)
        if self.shard.size > 1:
            self.cand_all = torch.zeros(self.rec_bytes * self.shard.size, dtype=torch.uint8,
                                        device=dev)
        else:
            self.cand_all = self.cand_out
        self.center_list = None
        self._state_host = torch.empty(64, dtype=torch.uint8).pin_memory()
        self.launches = 0
        self.center_store = self.center_store_traces = self.cc = None
        self._queued = 0            # host upper bound of the number of centres

    # -- fused peer-memory exchange --------------------------------------------------------
    def _setup_p2p(self):
        """Borrow (or create) the symmetric exchange buffer (torch symmetric memory: CUDA VMM +
        peer mapping) for the fused step + exchange kernels.  True when this rank is ready; the
        engine then agrees with its peers (ShardInfo.extras) and falls back to the NCCL
        all-gather unless every rank is."""
        import logging
        import os
        size = self.comm.size
        self._exch = None
        if size <= 1:
            return False
        want = (self.is_rmsd and not self.triangle and size <= 8
                and os.environ.get("ENSPARA_B200_P2P", "1") != "0")
        if not want:
            return False
        nbytes = int(self.lib.eb_exch_bytes(self.data.n_atoms, size))
        self._exch_key = (id(getattr(self.comm, "group", None)), size, nbytes, str(self.dev))
        item = _exchange_pool.take(self._exch_key)
        if item is None:
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm
                group = getattr(self.comm, "group", None) or dist.group.WORLD
                buf = symm.empty(nbytes, dtype=torch.uint8, device=self.dev)
                hdl = symm.rendezvous(buf, group)
                buf.zero_()
                peers = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64,
                                     device=self.dev)
                torch.cuda.synchronize()
                dist.barrier(group)   # every buffer is zeroed before anyone publishes into it
                item = (buf, hdl, peers)
            except Exception as exc:  # no peer access, old torch, ...
                logging.getLogger(__name__).warning(
                    "peer-memory candidate exchange unavailable (%r); using the NCCL "
                    "all-gather", exc)
                return False
        self._exch = item
        self._exch_buf, self._exch_hdl, self._exch_peers = item
        return True

    def _release_p2p(self, reusable):
        """Hand the exchange buffer back.  ``reusable`` only after a cleanly finished run: all
        ranks have passed the final collective, so nobody still reads or writes the buffer."""
        item, self._exch = self._exch, None
        if item is not None and reusable:
            _exchange_pool.give(self._exch_key, item)

    # -- state ---------------------------------------------------------------------------
    def set_state(self, distances, assignments):
        """Warm start (kcenters.py:200-206): take over existing distances / assignments."""
        self.dist.copy_(torch.as_tensor(np.asarray(distances), device=self.dev).to(
            self.dist.dtype))
        self.assign.copy_(torch.as_tensor(np.asarray(assignments), device=self.dev).to(
            torch.int32))

    def read_state(self):
        self._state_host.copy_(self.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        st = _lib.KcState.from_buffer_copy(self._state_host.numpy().tobytes())
        if st.error:
            self._release_p2p(reusable=False)
            raise RuntimeError(
                "enspara_b200: the peer-memory candidate exchange timed out on rank %d (a peer "
                "rank died or queued a different launch sequence); the run is void. "
                "ENSPARA_B200_P2P=0 selects the NCCL all-gather exchange." % self.shard.rank)
        return st

    # -- launches ------------------------------------------------------------------------
    def _exchange(self):
        if self.shard.size > 1:
            self.comm.all_gather_into(self.cand_all, self.cand_out)

    def seed(self, first_center_id=0):
        s = stream_ptr()
        if self.p2p and self._exch is None:      # a second run of the same engine
            self.p2p = self._setup_p2p()
            assert self.p2p, "peer-memory exchange could not be re-acquired"
        if self.p2p:
            d = self.data
            _lib.call("eb_kcenters_seed_rmsd_p2p", ptr(d.xyz), ptr(d.traces), self.n, d.n_atoms,
                      self.shard.offset, ptr(self._exch_peers), self.shard.size,
                      self.shard.rank, ptr(self.dist), int(first_center_id), ptr(self.state),
                      ptr(self.partials), ptr(self.cand_out), s)
            self.launches += 1
            return
        if self.is_rmsd:
            d = self.data
            _lib.call("eb_kcenters_seed_rmsd", ptr(d.xyz), ptr(d.traces), self.n, d.n_atoms,
                      self.shard.offset, ptr(self.dist), int(first_center_id), ptr(self.state),
                      ptr(self.partials), ptr(self.cand_out), s)
        else:
            d = self.data
            _lib.call("eb_kcenters_seed_feat", ptr(d.X), self.n, d.n_features, d.dt,
                      self.shard.offset, ptr(self.dist), int(first_center_id), ptr(self.state),
                      ptr(self.partials), ptr(self.cand_out), s)
        self.launches += 1
        self._exchange()

    def step(self, n_clusters_limit, cutoff, n_steps=1):
        """Queue ``n_steps`` iterations.  A single shard queues them inside one C call (no
        Python per launch); sharded runs interleave the candidate all-gather."""
        if self.p2p:
            d = self.data
            _lib.call("eb_kcenters_step_rmsd_p2p", ptr(d.xyz), ptr(d.traces), self.n, d.n_atoms,
                      self.shard.offset, ptr(self._exch_peers), self.shard.size,
                      self.shard.rank, ptr(self.dist), ptr(self.assign), n_clusters_limit,
                      float(cutoff), ptr(self.state), ptr(self.center_list), ptr(self.partials),
                      ptr(self.cand_out), int(self.exact), int(n_steps), stream_ptr())
            self.launches += n_steps
            self._queued += n_steps
            return
        if self.shard.size > 1 and n_steps > 1:
            for _ in range(n_steps):
                self.step(n_clusters_limit, cutoff, 1)
            return
        s = stream_ptr()
        d = self.data
        if self.is_rmsd and self.triangle:
            cap = int(self.center_store.shape[0])
            _lib.call("eb_kcenters_step_rmsd_tri", ptr(d.xyz), ptr(d.traces), self.n, d.n_atoms,
                      self.shard.offset, ptr(self.cand_all), self.shard.size, ptr(self.dist),
                      ptr(self.assign), n_clusters_limit, float(cutoff), ptr(self.state),
                      ptr(self.center_list), ptr(self.partials), ptr(self.cand_out),
                      ptr(self.center_store), ptr(self.center_store_traces), ptr(self.cc), cap,
                      int(self._queued), int(n_steps), s)
            self.launches += n_steps       # plus n_steps small centre-centre launches
        elif self.is_rmsd:
            _lib.call("eb_kcenters_step_rmsd", ptr(d.xyz), ptr(d.traces), self.n, d.n_atoms,
                      self.shard.offset, ptr(self.cand_all), self.shard.size, ptr(self.dist),
                      ptr(self.assign), n_clusters_limit, float(cutoff), ptr(self.state),
                      ptr(self.center_list), ptr(self.partials), ptr(self.cand_out),
                      int(self.exact), int(n_steps), s)
        else:
            _lib.call("eb_kcenters_step_feat", ptr(d.X), self.n, d.n_features, d.dt,
                      _lib_metric(self.kind), self.shard.offset, ptr(self.cand_all),
                      self.shard.size, ptr(self.dist), ptr(self.assign), n_clusters_limit,
                      float(cutoff), ptr(self.state), ptr(self.center_list),
                      ptr(self.partials), ptr(self.cand_out), int(n_steps), s)
        if not (self.is_rmsd and self.triangle):
            self.launches += n_steps
        self._queued += n_steps
        self._exchange()

    def _ensure_center_list(self, capacity):
        if self.center_list is None or self.center_list.numel() < capacity:
            new = torch.full((capacity,), -1, dtype=torch.int64, device=self.dev)
            if self.center_list is not None:
                new[:self.center_list.numel()].copy_(self.center_list)
            self.center_list = new
        if self.triangle and (self.center_store is None
                              or self.center_store.shape[0] < capacity):
            d = self.data
            store = torch.zeros((capacity, 3, d.a_pad), dtype=torch.float32, device=self.dev)
            tr = torch.zeros((capacity,), dtype=torch.float64, device=self.dev)
            cc = torch.zeros((capacity,), dtype=torch.float32, device=self.dev)
            if self.center_store is not None:
                m = int(self.center_store.shape[0])
                store[:m].copy_(self.center_store)
                tr[:m].copy_(self.center_store_traces)
            self.center_store, self.center_store_traces, self.cc = store, tr, cc

    def preload_centers(self, centers_dev):
        """Warm start in triangle mode: the existing centres (a DeviceTrajectory, in centre-id
        order) go into the store the pruning test reads."""
        if not self.triangle:
            return
        m = len(centers_dev)
        self._ensure_center_list(max(m, 1))
        self.center_store[:m].copy_(centers_dev.xyz[:m])
        self.center_store_traces[:m].copy_(centers_dev.traces[:m])

    def run(self, n_clusters, dist_cutoff, n_existing=0, on_progress=None):
        """Run until the reference's stop rule fires (kcenters.py:217).  ``n_existing`` centres
        (init_centers) already count towards ``n_clusters`` and shift the new centre ids.
        Returns (global frame indices of the centres chosen by this run, final global max of
        min-distances)."""
        n_global = self.shard.n_global
        if n_clusters is None or n_clusters == np.inf:
            limit = _INT32_MAX
        else:
            limit = int(min(int(n_clusters), _INT32_MAX))
        cutoff = 0.0 if dist_cutoff is None else float(dist_cutoff)
        self.seed(n_existing)
        if limit <= n_existing or n_global == 0:
            return [], self._finish()
        bounded = limit < _INT32_MAX
        # the centre list grows geometrically when only a cutoff bounds the run; a run can
        # never choose more new centres than there are frames
        hard_cap = n_existing + n_global
        capacity = min(limit, hard_cap) if bounded else min(hard_cap, n_existing + 4096)
        self._ensure_center_list(capacity)
        self._queued = n_existing
        queued = n_existing          # centre ids for which a launch has been queued
        poll = self.POLL_EVERY
        while True:
            if bounded and cutoff <= 0.0:
                batch = limit - queued
            else:
                # polls get rarer as the run gets longer (32, 64, 128, 128, ...): fewer state
                # reads, and batches of >= 64 launches tell the library that this is a long,
                # power-capped run (it then picks the kernel form that is faster sustained)
                batch = min(poll, limit - queued)
                poll = min(2 * poll, 4 * self.POLL_EVERY)
            if queued + batch > self.center_list.numel():
                self._ensure_center_list(min(hard_cap, max(2 * self.center_list.numel(),
                                                           queued + batch)))
                batch = min(batch, self.center_list.numel() - queued)
            if batch > 0:
                self.step(limit, cutoff, batch)
            queued += batch
            st = self.read_state()
            if on_progress is not None:
                on_progress(st)
            if st.done or st.n_centers >= limit or queued >= hard_cap or batch <= 0:
                break
        k = int(st.n_centers)
        centers = self.center_list[n_existing:k].cpu().numpy().astype(np.int64)
        self.wait_ns = int(st.wait_ns)
        return [int(c) for c in centers], self._finish()

    def _finish(self):
        """End of a run: the global max-min-distance (a collective every rank passes only after
        every rank's last step has completed), after which the exchange buffer is idle on all
        ranks and can go back to the pool."""
        m = self._global_maxdist()
        self._release_p2p(reusable=True)
        return m

    def _global_maxdist(self):
        st = self.read_state()
        t = torch.tensor([st.local_maxdist], dtype=torch.float64, device=self.dev)
        self.comm.all_reduce_max(t)
        return float(t.cpu()[0])

    # -- results -------------------------------------------------------------------------
    def results_host(self):
        """(assignments int64, distances float64) like kcenters.py:198-199."""
        return (self.assign.cpu().numpy().astype(np.int64),
                self.dist.cpu().numpy().astype(np.float64))


def _lib_metric(kind):
    return {"euclidean": _lib.METRIC_EUCLIDEAN, "manhattan": _lib.METRIC_MANHATTAN,
            "sqeuclidean": _lib.METRIC_SQEUCLIDEAN}[kind]

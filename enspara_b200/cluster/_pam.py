"""Device-resident PAM (k-medoids) sweeps.

Reference: /root/reference/enspara/cluster/kmedoids.py:520-699 (_kmedoids_pam_update) and
:482-517 (_propose_new_center_amongst).  Distances, assignments, the trial copies of both and
the medoid coordinates stay in HBM for the whole run; per proposal the host only draws the
random number (the reference's legacy ``RandomState`` stream must be consumed identically,
SURVEY.md App. A.5) and reads back three scalars (proposed index, number of frames needing a
full re-assignment, trial cost).

Sharded runs (one process per GPU, contiguous frame blocks): member counts are all-gathered,
every rank draws the same number from an identically seeded RandomState, the owner of the
selected member publishes the frame with one broadcast, every rank does its local pass, and the
trial cost is one float64 all-reduce -- so every rank takes the same accept decision.  The k-th
member is counted in GLOBAL frame order, which makes a sharded run equal to the serial one
(the reference's ``mpi.ops.randind`` maps the draw through a striped concatenation instead,
mpi/ops.py:256-268; that alternative map is available as ``striped_randind=True``).
"""
import time

import numpy as np
import torch
from sklearn.utils import check_random_state

from .. import _lib
from ..device import DeviceFeatures, DeviceTrajectory, ptr, stream_ptr
from ..exception import DataInvalid
from . import _ops
from ._engine import ShardInfo


class PamEngine:
    #: (frame, medoid) pairs below which the exact kernel is used for the ambiguous subset.
    #: The exact kernel walks all medoids serially per frame (3 us per medoid), so with >= 64
    #: medoids even a one-frame subset is faster through the screen (~0.1 ms fixed cost).
    TC_MIN_PAIRS = 0

    def __init__(self, data, metric, comm, distances, assignments, medoid_global_inds):
        self.data = data
        self.metric = metric
        self.comm = comm
        self.lib = _lib.load()
        self.n = len(data)
        self.is_rmsd = metric.is_rmsd
        self.dev = data.xyz.device if self.is_rmsd else data.X.device
        self.shard = ShardInfo(self.n, comm)
        self.ddt = torch.float32 if self.is_rmsd else torch.float64
        dev = self.dev
        self.dist = self._to_dev(distances, self.ddt)
        self.assign = self._to_dev(assignments, torch.int32)
        self.new_dist = torch.empty_like(self.dist)
        self.new_assign = torch.empty_like(self.assign)
        self.new_ctr_dist = torch.empty_like(self.dist)
        self.ambig_idx = torch.empty(max(self.n, 1), dtype=torch.int64, device=dev)
        # scalars read back per proposal: [proposed local index, n_ambig] and [cost numerator]
        # (one 32-byte block {cost f64, proposal i64, n_ambig i64, overflow i32} on the device
        # and in pinned host memory, so the read-back is a single copy)
        # followed by the need-list counter, so the three device counters of a proposal
        # {n_ambig, overflow, n_need} are adjacent and cleared by one memset
        self._scal_all = torch.zeros(6, dtype=torch.int64, device=dev)
        self.scal_d = self._scal_all[0:1].view(torch.float64)
        self.scal_i = self._scal_all[1:3]
        self._ovf = self._scal_all[3:4].view(torch.int32)[0:1]
        # zero-initialised once: eb_sum_squares keeps a ticket counter behind its partials
        self.scratch = torch.zeros(int(self.lib.eb_pam_scratch_bytes(self.n)),
                                   dtype=torch.uint8, device=dev)
        self.k = len(medoid_global_inds)
        self.counts = torch.zeros(max(self.k, 1), dtype=torch.int64, device=dev)
        self.medoid_global = [int(g) for g in medoid_global_inds]
        self.medoids = self._fetch_frames(self.medoid_global)
        self.prop, self.prop_buf, self.prop_idx = self._proposal_slot()
        self.saved = self._fetch_frames([])  # 1-frame scratch: the medoid a proposal displaces
        self.cost_num = self._sumsq(self.dist)
        self.n_global = self.shard.n_global
        # the ambiguous-subset re-assignment goes through the tcgen05 screen when it applies
        self.use_tc = self.is_rmsd and _ops.tc_applicable(metric, data, self.k)
        self._tc_ws = {}
        self.profile = None
        # triangle-inequality pruning of the full pass (RMSD is a metric): frames that provably
        # stay with their medoid are not read; the sweep's result is unchanged
        self.prune = self.is_rmsd and self.k >= 2
        self.cc = torch.zeros(max(self.k, 1), dtype=torch.float32, device=dev)
        #: compact-list form of the pruned pass (default) vs the fused skip-rounds kernel
        self.prune_compact = True
        if self.prune:
            self.need_idx = torch.empty(max(self.n, 1), dtype=torch.int64, device=dev)
            self.need_n = self._scal_all[4:5]
            self.need_assign = torch.zeros(max(self.n, 1), dtype=torch.int32, device=dev)
        self.counts_by_rank = None
        self._proposals_done = 0
        # pinned landing zone of the one read-back per proposal: [cost], [proposal, n_ambig],
        # [overflowed frames]
        self._pin_all = torch.zeros(4, dtype=torch.int64).pin_memory()
        self._pin_d = self._pin_all[0:1].view(torch.float64)
        self._pin_i = self._pin_all[1:3]
        self._pin_o = self._pin_all[3:4].view(torch.int32)[0:1]
        # zero-copy numpy views of the same pinned block: reading a scalar through torch
        # indexing costs microseconds per element, through numpy ~0.1 us
        self._np_i64 = self._pin_all.numpy()
        self._np_f64 = self._pin_all.view(torch.float64).numpy()
        self._np_i32 = self._pin_all.view(torch.int32).numpy()
        self.host_issue_s = 0.0
        self.sync_wait_s = 0.0
        self.refresh_s = 0.0
        self._t_proposal = time.perf_counter()
        # RMSD with pruning: the whole proposal is queued by ONE C call (eb_pam_propose_rmsd);
        # the per-step Python path below stays for feature metrics, profiling and the
        # synchronous audited proposals
        self._ctx = self._make_ctx() if (self.prune and self.prune_compact) else None
        self._fast_ws = {"cap": 0}
        self._red2 = torch.zeros(2, dtype=torch.float64, device=dev)
        self._pin_d2 = torch.zeros(2, dtype=torch.float64).pin_memory()

    # -- helpers -------------------------------------------------------------------------
    def _to_dev(self, arr, dtype):
        if torch.is_tensor(arr):
            return arr.to(device=self.dev, dtype=dtype).clone()
        a = np.ascontiguousarray(np.asarray(arr))
        if len(a) != self.n:
            raise DataInvalid("length mismatch: %d values for %d frames" % (len(a), self.n))
        return torch.from_numpy(a).to(self.dev).to(dtype)

    def _empty_frames(self, m):
        if self.is_rmsd:
            return DeviceTrajectory.empty(m, self.data.n_atoms, self.data.top)
        return DeviceFeatures(torch.zeros((m, self.data.n_features), dtype=self.data.X.dtype,
                                          device=self.dev))

    def _proposal_slot(self):
        """One flat buffer [frame payload | trace (f64) | global index (i64)] so that a sharded
        proposal travels in ONE broadcast; returns (1-frame container viewing the payload,
        the flat buffer, the int64 view of the index)."""
        if self.is_rmsd:
            a_pad = self.data.a_pad
            nf = 3 * a_pad
            buf = torch.zeros(nf + 4, dtype=torch.float32, device=self.dev)
            xyz = buf[:nf].view(1, 3, a_pad)
            traces = buf[nf:nf + 2].view(torch.float64)
            idx = buf[nf + 2:nf + 4].view(torch.int64)
            return DeviceTrajectory(xyz, traces, self.data.n_atoms, self.data.top), buf, idx
        row_bytes = self.data.n_features * self.data.X.element_size()
        pad = (row_bytes + 7) // 8 * 8
        buf = torch.zeros(pad + 8, dtype=torch.uint8, device=self.dev)
        X = buf[:row_bytes].view(self.data.X.dtype).view(1, self.data.n_features)
        idx = buf[pad:pad + 8].view(torch.int64)
        return DeviceFeatures(X), buf, idx

    def _fetch_frames(self, global_inds):
        """Dense device copy of the frames with these GLOBAL indices on every rank."""
        m = len(global_inds)
        out = self._empty_frames(max(m, 1))
        if m == 0:
            return out
        sh = self.shard
        mine = [(j, g - sh.offset) for j, g in enumerate(global_inds)
                if sh.offset <= g < sh.offset + self.n]
        if sh.size > 1:
            if self.is_rmsd:
                out.xyz.zero_()
                out.traces.zero_()
            else:
                out.X.zero_()
        if mine:
            loc = torch.as_tensor([l for _, l in mine], dtype=torch.int64, device=self.dev)
            pos = torch.as_tensor([j for j, _ in mine], dtype=torch.int64, device=self.dev)
            if self.is_rmsd:
                sub = self.data.gather(loc)
                out.xyz[pos] = sub.xyz
                out.traces[pos] = sub.traces
            else:
                out.X[pos] = self.data.X[loc]
        if sh.size > 1:
            if self.is_rmsd:
                self.comm.all_reduce_sum(out.xyz)
                self.comm.all_reduce_sum(out.traces)
            else:
                self.comm.all_reduce_sum(out.X)
        return out

    def _sumsq(self, vec):
        _lib.call("eb_sum_squares", ptr(vec), self.n, int(not self.is_rmsd), ptr(self.scal_d),
                  ptr(self.scratch), stream_ptr())
        t = self.scal_d.clone()
        self.comm.all_reduce_sum(t)
        return float(t.cpu()[0])

    def _sumsq_with(self, vec, scal_i, ovf):
        """Global sum of squares of ``vec`` plus the proposal's integer scalars and the screen's
        overflow counter, all fetched with ONE stream synchronisation."""
        _lib.call("eb_sum_squares", ptr(vec), self.n, int(not self.is_rmsd), ptr(self.scal_d),
                  ptr(self.scratch), stream_ptr())
        t = self.scal_d
        if self.shard.size > 1:
            t = self.scal_d.clone()
            self.comm.all_reduce_sum(t)
        self._pin_d.copy_(t, non_blocking=True)
        self._pin_i.copy_(scal_i, non_blocking=True)
        if ovf is not None:
            self._pin_o.copy_(ovf, non_blocking=True)
        t0 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t1 = time.perf_counter()
        # diagnostic split of a proposal's wall time: host issuing work vs waiting for the GPU
        self.host_issue_s += t0 - self._t_proposal
        self.sync_wait_s += t1 - t0
        return (float(self._pin_d[0]), self._pin_i.clone(),
                int(self._pin_o[0]) if ovf is not None else 0)

    def _refresh_counts(self):
        """Member counts of every cluster on every rank (len(np.where(assignments == cid)[0]),
        kmedoids.py:611).  Assignments only change when a proposal is accepted, so the counts
        are cached on the host between acceptances (one histogram launch + one small D2H)."""
        t0 = time.perf_counter()
        self._refresh_counts_impl()
        self.refresh_s += time.perf_counter() - t0

    def _refresh_counts_impl(self):
        _lib.call("eb_count_members", ptr(self.assign), self.n, self.k, ptr(self.counts),
                  stream_ptr())
        if self.shard.size > 1:
            allc = torch.empty(self.shard.size * self.k, dtype=torch.int64, device=self.dev)
            self.comm.all_gather_into(allc, self.counts[:self.k].contiguous())
            self.counts_by_rank = allc.view(self.shard.size, self.k).cpu().numpy()
        else:
            self.counts_by_rank = self.counts[:self.k].cpu().numpy().astype(np.int64)[None]

    def _member_counts(self, cid):
        return self.counts_by_rank[:, cid]

    def _slot_copy(self, dst, j, src, i):
        if self.is_rmsd:
            dst.xyz[j].copy_(src.xyz[i])
            dst.traces[j].copy_(src.traces[i])
        else:
            dst.X[j].copy_(src.X[i])

    def _load_proposal(self, owner, local_idx_dev=None, local_idx=None):
        """Make self.prop (1 frame) and self.prop_idx (its GLOBAL index) valid on every rank:
        the owner gathers the frame, one broadcast of the flat buffer does the rest
        (the reference: bcast of the index + Bcast of the frame, kmedoids.py:500-508)."""
        sh = self.shard
        if sh.rank == owner:
            idx = local_idx_dev if local_idx_dev is not None else torch.as_tensor(
                [int(local_idx)], dtype=torch.int64, device=self.dev)
            if self.is_rmsd:
                _lib.call("eb_gather_frames", ptr(self.data.xyz), ptr(self.data.traces),
                          self.data.n_atoms, ptr(idx), 1, ptr(self.prop.xyz),
                          ptr(self.prop.traces), stream_ptr())
            else:
                self.prop.X[0].copy_(self.data.X[idx[0]])
            torch.add(idx[:1], sh.offset, out=self.prop_idx)
        if sh.size > 1:
            self.comm.broadcast(self.prop_buf, owner)

    # -- one-call proposals (RMSD) ---------------------------------------------------------
    def _make_ctx(self):
        import ctypes
        c = _lib.PamCtx()
        d = self.data
        c.xyz, c.traces, c.n, c.frame_offset = ptr(d.xyz), ptr(d.traces), self.n, self.shard.offset
        c.n_atoms, c.k = d.n_atoms, self.k
        c.medoid_xyz, c.medoid_traces = ptr(self.medoids.xyz), ptr(self.medoids.traces)
        c.prop_xyz, c.prop_traces, c.prop_idx = (ptr(self.prop.xyz), ptr(self.prop.traces),
                                                 ptr(self.prop_idx))
        c.saved_xyz, c.saved_traces = ptr(self.saved.xyz), ptr(self.saved.traces)
        c.new_ctr_dist, c.cc = ptr(self.new_ctr_dist), ptr(self.cc)
        c.need_idx, c.need_n, c.need_assign = (ptr(self.need_idx), ptr(self.need_n),
                                               ptr(self.need_assign))
        c.ambig_idx, c.scal_i, c.scal_d = ptr(self.ambig_idx), ptr(self.scal_i), ptr(self.scal_d)
        c.scratch = ptr(self.scratch)
        c.kappa = float(_ops.tc_kappa(d.a_pad))
        c.pin_d, c.pin_i, c.pin_o = (self._pin_d.data_ptr(), self._pin_i.data_ptr(),
                                     self._pin_o.data_ptr())
        # medoids the triangle inequality leaves for the ambiguous frames (PAM):
        # ENSPARA_B200_PAM_LIST=0 keeps the all-medoids re-assignment
        import os
        cap = min(self.k, 1024)
        self._med_list = torch.zeros(cap + 2, dtype=torch.int32, device=self.dev)
        c.med_list, c.med_list_n, c.med_list_cap = (ptr(self._med_list[2:]),
                                                    ptr(self._med_list[:1]), cap)
        c.use_list = int(os.environ.get("ENSPARA_B200_PAM_LIST", "1") != "0")
        self._list_overflows = 0
        self._ctx_ref = ctypes.byref(c)
        self._ctx = c
        self._ctx_state()
        fn, self._restore = self.lib.eb_pam_propose_rmsd, self.lib.eb_pam_restore_medoid
        check = _lib.check
        self._propose = lambda *a: check(fn(*a))
        return c

    def _ctx_state(self):
        """Current / trial state pointers (they swap when a proposal is accepted)."""
        c = self._ctx
        c.dist, c.assign = ptr(self.dist), ptr(self.assign)
        c.new_dist, c.new_assign = ptr(self.new_dist), ptr(self.new_assign)

    def _ctx_workspace(self, m_max):
        """Screen workspace of the one-call path, grown geometrically with the member count."""
        ws = self._fast_ws
        if ws.get("ovf") is None:
            ws["ovf"] = self._ovf
            self._ctx.tc_ovf = ptr(ws["ovf"])
        if self._ctx.use_list:
            return True                  # overflow counter: more listed medoids than the cap
        use_tc = bool(self.use_tc and m_max * self.k >= self.TC_MIN_PAIRS
                      and 0 < m_max <= _ops.TC_CHUNK_FRAMES)
        if use_tc and m_max > ws["cap"]:
            cap = min(_ops.TC_CHUNK_FRAMES, m_max + m_max // 4 + 128)
            need = int(self.lib.eb_tc_scratch_bytes(cap, self.data.n_atoms, self.k))
            ws["scratch"] = ws["cand"] = None
            ws["scratch"] = torch.empty(need, dtype=torch.uint8, device=self.dev)
            ws["cand"] = torch.empty(cap, dtype=torch.int32, device=self.dev)
            ws["cap"] = cap
            self._ctx.tc_scratch, self._ctx.tc_cand = ptr(ws["scratch"]), ptr(ws["cand"])
        self._ctx.use_tc = int(use_tc)
        return use_tc

    def _readback(self, with_ovf):
        """The proposal's scalars with ONE stream synchronisation (sharded: after the cost's
        all-reduce); the single-GPU one-call path has queued the copies already."""
        sharded = self.shard.size > 1
        if sharded:
            # cost and overflow count travel in one all-reduce: every rank must take the same
            # decision about the exact fallback (its cost is a collective again)
            t = self._red2
            t[0:1].copy_(self.scal_d)
            t[1:2].copy_(self._fast_ws["ovf"])
            self.comm.all_reduce_sum(t)
            self._pin_d2.copy_(t, non_blocking=True)
            self._pin_i.copy_(self.scal_i, non_blocking=True)
        t0 = time.perf_counter()
        self._stream.synchronize()
        t1 = time.perf_counter()
        self.host_issue_s += t0 - self._t_proposal
        self.sync_wait_s += t1 - t0
        scal = (int(self._np_i64[1]), int(self._np_i64[2]))
        if sharded:
            return float(self._pin_d2[0]), scal, int(self._pin_d2[1])
        return float(self._np_f64[0]), scal, (int(self._np_i32[6]) if with_ovf else 0)

    def _draw_member(self, cid, rs, striped_randind):
        """(owner rank, k-th member on the owner) of a random member of cluster cid, consuming
        the reference's random stream (kmedoids.py:482-517, mpi/ops.py:247-268)."""
        sh = self.shard
        n_states = self._member_counts(cid)
        total = int(n_states.sum())
        if total < 1:
            raise ValueError("'a' cannot be empty unless no samples are taken")
        g = int(rs.randint(total))
        if sh.size == 1:
            return 0, g
        if striped_randind:
            concat = np.concatenate([np.arange(total)[r::sh.size] for r in range(sh.size)])
            g = int(np.where(concat == g)[0][0])
        bounds = np.concatenate([[0], np.cumsum(n_states)])
        owner = int(np.searchsorted(bounds, g, side="right") - 1)
        return owner, g - int(bounds[owner])

    def _fast_proposal(self, cid, proposals, rs, striped_randind):
        """One proposal through eb_pam_propose_rmsd.  Returns ((accepted, old_cost, new_cost,
        accepted), proposal's global index) -- the tuple's tail is what the sweep logs."""
        sh = self.shard
        ctx, stream = self._ctx_ref, self._stream_ptr
        call = self._propose
        m_max = int(self.counts_by_rank[sh.rank, cid])
        use_tc = self._ctx_workspace(m_max)
        stages = _lib.PAM_TRIAL
        kth = 0
        if proposals is None:
            owner, kth = self._draw_member(cid, rs, striped_randind)
            prop_global = None
            if sh.size == 1:
                stages |= _lib.PAM_SELECT
            else:
                if sh.rank == owner:
                    call(ctx, cid, kth, 0, _lib.PAM_SELECT, stream)
                self.comm.broadcast(self.prop_buf, owner)
        else:
            prop_global = int(proposals[cid])
            owner, loc = sh.to_rank_local(prop_global)
            self._load_proposal(owner, local_idx=loc)
        if sh.size == 1:
            stages |= _lib.PAM_READBACK
        call(ctx, cid, kth, m_max, stages, stream)
        new_num, scal, n_ovf = self._readback(use_tc)
        if n_ovf > 0:
            # a candidate list (screen) or the medoid list overflowed somewhere: the subset
            # goes through the general path on every rank, and the cost is taken again
            n_ambig = int(scal[1])
            if n_ambig > 0 and self.use_tc:
                _ops.assign_device_tc(self.metric, self.data, self.medoids, k=self.k,
                                      frame_idx=self.ambig_idx, n_idx=n_ambig,
                                      out_dist=self.new_dist, out_assign=self.new_assign,
                                      scatter=True, workspace=self._tc_ws)
            else:
                _ops.assign_device(self.metric, self.data, self.medoids,
                                   frame_idx=self.ambig_idx, n_idx=n_ambig,
                                   out_dist=self.new_dist, out_assign=self.new_assign,
                                   accumulate=False, scatter=True, k=self.k)
            new_num = self._sumsq(self.new_dist)
            if self._ctx.use_list:
                # data without cluster structure: after a few overflows stop listing
                self._list_overflows += 1
                if self._list_overflows >= 4:
                    self._ctx.use_list = 0
        if prop_global is None:
            prop_global = int(scal[0])
        old_cost = self.cost_num / self.n_global
        new_cost = new_num / self.n_global
        accepted = new_cost < old_cost
        if accepted:
            self.dist, self.new_dist = self.new_dist, self.dist
            self.assign, self.new_assign = self.new_assign, self.assign
            self._ctx_state()
            self.cost_num = new_num
            self.medoid_global[cid] = prop_global
        else:
            _lib.check(self._restore(ctx, cid, stream))
        return (bool(accepted), old_cost, new_cost, bool(accepted)), prop_global

    # -- one sweep -----------------------------------------------------------------------
    def sweep(self, proposals=None, random_state=None, striped_randind=False, log=None,
              max_proposals=None):
        """One pass over all k clusters (kmedoids.py:609-694).  ``proposals``: GLOBAL frame
        indices, one per cluster, or None for random proposals.  Returns acceptances.
        ``max_proposals`` stops after that many clusters (benchmarks only)."""
        rs = check_random_state(random_state)
        sh = self.shard
        if sh.size > 1 and proposals is None:
            # Every rank derives owner / k-th member / broadcast root from its own draw, so all
            # ranks must consume ONE random stream: rank 0's (the reference: rank 0 draws and
            # broadcasts the index, mpi/ops.py:247-253).  With random_state=None the ranks'
            # global RandomStates are seeded differently; one broadcast of rank 0's generator
            # state per sweep aligns them, and a seeded run still equals the serial run.
            state = self.comm.broadcast_object(rs.get_state() if sh.rank == 0 else None, 0)
            if sh.rank != 0:
                if rs is np.random.mtrand._rand:     # do not clobber the process-global state
                    rs = np.random.RandomState()
                rs.set_state(state)
        acceptances = 0
        prof = self.profile          # None, or dict phase -> seconds (developer timing)
        # the current stream does not change during a sweep: look it up once
        self._stream = torch.cuda.current_stream()
        self._stream_ptr = stream_ptr()

        def tick(name, t0):
            if prof is None:
                return 0.0
            torch.cuda.synchronize()
            now = time.perf_counter()
            if name:
                prof[name] = prof.get(name, 0.0) + now - t0
            return now
        t = tick(None, 0.0)
        self._refresh_counts()
        t = tick("counts", t)
        for cid in range(self.k if max_proposals is None else min(self.k, max_proposals)):
            # ---- proposal (kmedoids.py:616-628, 482-517) -------------------------------
            self._t_proposal = time.perf_counter()
            # Proposals are queued by ONE C call and cost one stream synchronisation.  While
            # the subset is re-assigned through the tensor-core screen (medoid list off or
            # given up), every TC_AUDIT_EVERY-th proposal takes the step-by-step path below,
            # whose synchronous screen call also runs the audit; so does any profiled one.
            fast = (self._ctx is not None and self.prune and self.prune_compact
                    and prof is None and self.counts_by_rank is not None
                    and (self._ctx.use_list == 1
                         or (self._proposals_done + 1) % _ops.TC_AUDIT_EVERY != 0
                         or _ops._audit_level() == 0))
            if fast:
                accepted, prop_global = self._fast_proposal(cid, proposals, rs, striped_randind)
                self._proposals_done += 1
                if log is not None:
                    log.append((cid, prop_global) + accepted[1:])
                if accepted[0]:
                    acceptances += 1
                    if cid + 1 < self.k:
                        self._refresh_counts()
                continue
            if proposals is None:
                owner, kth = self._draw_member(cid, rs, striped_randind)
                if sh.rank == owner:
                    _lib.call("eb_select_member", ptr(self.assign), self.n, cid, kth,
                              ptr(self.scal_i), ptr(self.scratch), stream_ptr())
                    self._load_proposal(owner, local_idx_dev=self.scal_i[:1])
                else:
                    self._load_proposal(owner)
                prop_global = None  # resolved after the read-back below
            else:
                prop_global = int(proposals[cid])
                owner, loc = sh.to_rank_local(prop_global)
                self._load_proposal(owner, local_idx=loc)

            t = tick("proposal", t)
            # ---- full pass + three-way split (kmedoids.py:637-658) -------------------
            if self.prune:
                # distances proposal -> every medoid (k evaluations), then the pruned pass
                _ops.one_to_all_device(self.metric, self.medoids, self.prop, out=self.cc)
                d = self.new_ctr_dist
                if self.prune_compact:
                    # frames that provably keep their medoid get +inf; the rest, as a compact
                    # list, are evaluated against the proposal by the exact kernel (k = 1) with
                    # every lane busy; same values as the fused pruned pass / a full pass
                    _lib.call("eb_pam_need_list", ptr(self.dist), ptr(self.assign), ptr(self.cc),
                              self.n, cid, ptr(d), ptr(self.need_idx), ptr(self.need_n),
                              stream_ptr())
                    _ops.assign_device(self.metric, self.data, self.prop,
                                       frame_idx=self.need_idx, n_idx=self.n, out_dist=d,
                                       out_assign=self.need_assign, accumulate=False,
                                       scatter=True, k=1, n_dev=self.need_n)
                else:
                    _lib.call("eb_rmsd_one_to_all_pruned", ptr(self.data.xyz),
                              ptr(self.data.traces), self.n, self.data.n_atoms,
                              ptr(self.prop.xyz), ptr(self.prop.traces), ptr(self.dist),
                              ptr(self.assign), ptr(self.cc), cid, ptr(d), stream_ptr())
            else:
                d = _ops.one_to_all_device(self.metric, self.data, self.prop,
                                           out=self.new_ctr_dist)
            t = tick("full_pass", t)
            _lib.call("eb_pam_classify", ptr(d), ptr(self.dist), ptr(self.assign), self.n,
                      int(not self.is_rmsd), cid, ptr(self.new_dist), ptr(self.new_assign),
                      ptr(self.ambig_idx), ptr(self.scal_i[1:]), stream_ptr())
            self.scal_i[:1].copy_(self.prop_idx)
            # ---- ambiguous frames against all medoids, proposal in slot cid (:660-670) --
            self._slot_copy(self.saved, 0, self.medoids, cid)
            self._slot_copy(self.medoids, cid, self.prop, 0)
            self._proposals_done += 1
            # RMSD: the size of the ambiguous subset stays on the device (every ambiguous
            # frame is a member of cluster cid, so the cached member count bounds it); the
            # host reads it back together with the trial cost -- ONE synchronisation per
            # proposal.  Every TC_AUDIT_EVERY-th proposal takes the synchronous path, which
            # also runs the screen's audit.
            deferred = (self.is_rmsd and self.counts_by_rank is not None
                        and (self._proposals_done % _ops.TC_AUDIT_EVERY != 0
                             or _ops._audit_level() == 0) and prof is None)
            ovf = None
            if deferred:
                m_max = int(self._member_counts(cid)[sh.rank])
                if m_max > 0:
                    if self.use_tc and m_max * self.k >= self.TC_MIN_PAIRS \
                            and m_max <= _ops.TC_CHUNK_FRAMES:
                        _, _, ovf = _ops.assign_device_tc(
                            self.metric, self.data, self.medoids, k=self.k,
                            frame_idx=self.ambig_idx, n_idx=m_max, out_dist=self.new_dist,
                            out_assign=self.new_assign, scatter=True, workspace=self._tc_ws,
                            n_dev=self.scal_i[1:], defer=True)
                    else:
                        _ops.assign_device(self.metric, self.data, self.medoids,
                                           frame_idx=self.ambig_idx, n_idx=m_max,
                                           out_dist=self.new_dist, out_assign=self.new_assign,
                                           accumulate=False, scatter=True, k=self.k,
                                           n_dev=self.scal_i[1:])
                new_num, scal, n_ovf = self._sumsq_with(self.new_dist, self.scal_i, ovf)
                n_ambig = int(scal[1])
                if n_ovf > 0:
                    # a candidate list overflowed: exact kernel for the subset, cost again
                    _ops.assign_device(self.metric, self.data, self.medoids,
                                       frame_idx=self.ambig_idx, n_idx=n_ambig,
                                       out_dist=self.new_dist, out_assign=self.new_assign,
                                       accumulate=False, scatter=True, k=self.k)
                    new_num = self._sumsq(self.new_dist)
            else:
                scal = self.scal_i.cpu()
                n_ambig = int(scal[1])
                t = tick("classify+readback", t)
                if n_ambig > 0:
                    if self.use_tc and n_ambig * self.k >= self.TC_MIN_PAIRS:
                        # tensor-core screen + exact re-score: same result as the exact kernel
                        _ops.assign_device_tc(self.metric, self.data, self.medoids, k=self.k,
                                              frame_idx=self.ambig_idx, n_idx=n_ambig,
                                              out_dist=self.new_dist,
                                              out_assign=self.new_assign, scatter=True,
                                              workspace=self._tc_ws)
                    else:
                        _ops.assign_device(self.metric, self.data, self.medoids,
                                           frame_idx=self.ambig_idx, n_idx=n_ambig,
                                           out_dist=self.new_dist, out_assign=self.new_assign,
                                           accumulate=False, scatter=True, k=self.k)
                t = tick("subset_assign", t)
                new_num = self._sumsq(self.new_dist)
            if prop_global is None:
                prop_global = int(scal[0])
            if prof is not None:
                prof["n_ambig"] = prof.get("n_ambig", 0) + n_ambig
            # ---- accept / reject on mean-square cost (kmedoids.py:680-694) -----------------
            old_cost = self.cost_num / self.n_global
            new_cost = new_num / self.n_global
            accepted = new_cost < old_cost
            if log is not None:
                log.append((cid, prop_global, old_cost, new_cost, bool(accepted)))
            if accepted:
                self.dist, self.new_dist = self.new_dist, self.dist
                self.assign, self.new_assign = self.new_assign, self.assign
                if self._ctx is not None:
                    self._ctx_state()
                self.cost_num = new_num
                self.medoid_global[cid] = prop_global
                acceptances += 1
                if cid + 1 < self.k:
                    self._refresh_counts()
            else:
                self._slot_copy(self.medoids, cid, self.saved, 0)
            t = tick("cost+accept", t)
        self.last_cost = self.cost_num / self.n_global
        return acceptances

    def results_host(self):
        return (self.assign.cpu().numpy().astype(np.int64),
                self.dist.cpu().numpy().astype(np.float64))

/*
 * enspara_b200.h -- C ABI of the B200-native clustering hot path (libenspara_b200.so).
 *
 * This is the drop-in boundary for enspara's conformational-clustering path.  The reference
 * is Python that reaches native code through two FFI seams; each entry point below names the
 * reference interface it replaces (paths under /root/reference/enspara/):
 *
 *   seam 1  mdtraj.rmsd (C/SSE, third party)           cluster/util.py:290-291, apps/cluster.py:210
 *   seam 2  geometry/libdist.pyx (Cython/OpenMP)        libdist.pyx:100-183
 *   plus the numpy bookkeeping wrapped around them      cluster/kcenters.py:282-309,
 *                                                       cluster/util.py:186-203,
 *                                                       cluster/kmedoids.py:611-694
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), never
 *     allocates device memory and never synchronises, unless stated;
 *   - return value 0 = ok; non-zero = error, text via eb_last_error() (thread local);
 *     EB_ERR_INVALID maps to enspara.exception.DataInvalid in the Python host layer;
 *   - frames of a trajectory live as float32 "SoA blocks": frame f occupies 3*A_pad floats,
 *     x[0..A_pad) y[0..A_pad) z[0..A_pad), A_pad = eb_rmsd_apad(A) (multiple of 8 so every
 *     row starts on a 32-byte sector), padding atoms are 0; coordinates are pre-centred and
 *     traces[f] = sum |x|^2 (float64) is stored beside them.
 *   - a "candidate record" is how one shard publishes its farthest frame after a step:
 *       [ 0] double  dist      current max of min-distances on the shard (-1 if shard empty)
 *       [ 8] int64   index     GLOBAL frame index of that frame (lowest index among ties)
 *       [16] double  trace     RMSD: trace of that frame; else 0
 *       [24] int64   reserved
 *       [32] payload           RMSD: 3*A_pad float32 SoA coords; euclid: F elements of dtype
 *     Records are what ranks exchange (one NCCL all-gather per k-centers iteration); the next
 *     step picks the winner among the gathered records on the device, so the host never needs
 *     to know which rank owns the new centre.
 */
#ifndef ENSPARA_B200_H
#define ENSPARA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EB_OK 0
#define EB_ERR_INVALID 1 /* bad shapes / arguments  -> DataInvalid            */
#define EB_ERR_CUDA 2    /* CUDA runtime error      -> RuntimeError           */
#define EB_ERR_LIMIT 3   /* size beyond a documented limit -> RuntimeError    */

/* element types accepted by the libdist replacements (libdist.pyx:9-15 FLOAT_TYPE_T) */
#define EB_DT_F32 0
#define EB_DT_F64 1
#define EB_DT_I8 2
#define EB_DT_I16 3
#define EB_DT_I32 4
#define EB_DT_I64 5

/* metrics of the fused feature-vector kernels (cluster/util.py:289-313) */
#define EB_METRIC_EUCLIDEAN 0
#define EB_METRIC_MANHATTAN 1
#define EB_METRIC_SQEUCLIDEAN 2 /* the squared-euclid callable of test_cluster.py:509-510 */

/* k-centers device state, one per shard; read back by the host once per batch of steps */
typedef struct eb_kc_state {
    int32_t n_centers;    /* centres chosen so far                                        */
    int32_t done;         /* 1 once the stop rule of kcenters.py:217 fired                */
    uint32_t blocks_done; /* internal: last-block ticket                                  */
    int32_t n_noop;       /* steps that were skipped because done==1 (diagnostic)         */
    double maxdist;       /* GLOBAL max of min-distances seen by the last step's prologue */
    double local_maxdist; /* max of min-distances on this shard after the last step       */
    int64_t last_center;  /* global index of the last centre chosen                       */
    int64_t error;        /* != 0: the peer-memory exchange timed out (a rank died or queued a
                             different launch sequence); the run is void, later steps no-ops */
    int64_t wait_ns;      /* fused exchange: ns block 0 spent waiting for the peers' records,
                             summed over the steps since the seed (diagnostic)               */
    int64_t reserved;
} eb_kc_state; /* 64 bytes */

int eb_version(void);
const char *eb_last_error(void);
/* number of SMs of the current device (grid sizing is done inside the library) */
int eb_sm_count(void);

/* ---- layout helpers ------------------------------------------------------------------- */
int eb_rmsd_apad(int n_atoms);
size_t eb_rmsd_record_bytes(int n_atoms);
/* bytes of scratch ("partials") the step / seed kernels need; constant per device */
size_t eb_kc_partials_bytes(void);

/* ---- K5: centring + trace, AoS -> SoA --------------------------------------------------
 * Replaces the per-call copy + inplace_center_and_trace_atom_major of mdtraj.rmsd (done once
 * here instead of once per distance call).  xyz_aos is (n, n_atoms, 3) float32 C-order as in
 * md.Trajectory.xyz.  precentered != 0 skips the centroid subtraction (cluster/util.py:625-629
 * `precentered=True`).  */
int eb_center_and_trace(const float *xyz_aos, int64_t n, int n_atoms, int precentered,
                        float *xyz_soa, double *traces, void *stream);
/* inverse layout change for handing centres back to the host as (n, n_atoms, 3) */
int eb_soa_to_aos(const float *xyz_soa, int64_t n, int n_atoms, float *xyz_aos, void *stream);
/* gather frames idx[0..m) (local indices) of an SoA block into a dense SoA block + traces */
int eb_gather_frames(const float *xyz_soa, const double *traces, int n_atoms,
                     const int64_t *idx, int64_t m, float *out_soa, double *out_traces,
                     void *stream);

/* ---- K1: fused k-centers step, RMSD -----------------------------------------------------
 * One launch == one iteration of kcenters.py:243-311 (and :314-378 for a shard):
 *   prologue  pick the winner among n_cand candidate records (max dist, lowest global index)
 *             -> new centre; apply the stop rule of kcenters.py:217
 *             (n_centers < n_clusters_limit && maxdist > dist_cutoff), else set state->done;
 *   body      stream every frame once, 3x3 inner-product matrix in float64, Theobald QCP,
 *             float32 RMSD, strict '<' update of dist/assign (kcenters.py:304-306);
 *   epilogue  shard arg-max (first occurrence) -> cand_out record, centre list append.
 * cand_in may alias cand_out (single GPU).  dist is float32 (values are float32 in the
 * reference too: md.rmsd returns float32), assign int32; the host layer widens them to the
 * reference's float64 / int64 when it copies results out.  exact=1 accumulates the matrix in
 * float64 (parity mode, default); exact=0 uses float32 blocks of 4 atoms + float64 block sums.
 * n_steps > 1 queues that many consecutive iterations in one call (single shard only: the
 * next step reads the record this one wrote; sharded runs interleave the all-gather). */
/* 1 when the exact step of a shard of n frames runs the TMA-staged kernel
 * (k_kcenters_step_rmsd_tma), 0 when it runs the LDG kernel; reporting aid for benchmarks. */
int eb_kcenters_step_rmsd_uses_tma(int64_t n, int n_atoms);
int eb_kcenters_step_rmsd(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          int64_t frame_offset, const void *cand_in, int n_cand,
                          float *dist, int32_t *assign, int32_t n_clusters_limit,
                          double dist_cutoff, eb_kc_state *state, int64_t *center_list,
                          void *partials, void *cand_out, int exact, int n_steps, void *stream);
/* seed: arg-max of the CURRENT dist array -> cand_out (all +inf gives frame 0, which is the
 * reference's first centre, kcenters.py:199,282,326-328); resets *state with
 * n_centers = first_center_id (> 0 when continuing from init_centers, kcenters.py:200-206). */
/* Fused step + candidate exchange over peer memory (NVLink P2P).  Replaces, per iteration, the
 * reference's two allgathers + Bcast + Barrier + allreduce (cluster/kcenters.py:332-348,
 * mpi/ops.py:137-138) WITHOUT any collective launch: the last block of a step stores its shard's
 * candidate record into every peer's symmetric buffer and releases a per-rank flag; the next
 * step's prologue acquires the flags of its own buffer.  peers_dev: device array of n_ranks
 * int64 base addresses of the ranks' exchange buffers (eb_exch_bytes(n_atoms, n_ranks) bytes
 * each, zero-initialised, mapped into every peer -- e.g. torch symmetric memory); every rank
 * must queue the same sequence of seed / step launches.  n_steps > 1 is allowed. */
size_t eb_exch_bytes(int n_atoms, int n_ranks);
int eb_kcenters_seed_rmsd_p2p(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *peers_dev, int n_ranks, int rank,
                              const float *dist, int32_t first_center_id, eb_kc_state *state,
                              void *partials, void *cand_out, void *stream);
int eb_kcenters_step_rmsd_p2p(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *peers_dev, int n_ranks, int rank,
                              float *dist, int32_t *assign, int32_t n_clusters_limit,
                              double dist_cutoff, eb_kc_state *state, int64_t *center_list,
                              void *partials, void *cand_out, int exact, int n_steps,
                              void *stream);
/* Triangle-inequality variant of the step (cluster/kcenters.py:287-296,
 * use_triangle_inequality=True).  center_store (store_capacity x 3*A_pad floats) and
 * center_store_traces hold the coordinates of the centres chosen so far (slots
 * [0, n_centers)); cc (store_capacity floats) receives d(new centre, centre j).  k_upper: a host
 * upper bound of n_centers before the first of the n_steps launches (sizes the small
 * centre-centre grid).  Results are identical to eb_kcenters_step_rmsd. */
int eb_kcenters_step_rmsd_tri(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *cand_in, int n_cand, float *dist,
                              int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                              eb_kc_state *state, int64_t *center_list, void *partials,
                              void *cand_out, float *center_store, double *center_store_traces,
                              float *cc, int64_t store_capacity, int64_t k_upper, int n_steps,
                              void *stream);
int eb_kcenters_seed_rmsd(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          int64_t frame_offset, const float *dist, int32_t first_center_id,
                          eb_kc_state *state, void *partials, void *cand_out, void *stream);
/* one-vs-all RMSD only (the md.rmsd(traj, frame) call itself, util.py:290): out float32[n];
 * center_trace points at the centre's trace in device memory */
int eb_rmsd_one_to_all(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                       const float *center_soa, const double *center_trace, float *out,
                       int exact, void *stream);

/* ---- K3: many-centres nearest-centre assignment (cluster/util.py:159-205) ---------------
 * For every frame the nearest of k centres, strict '<' in centre order (lowest centre index
 * wins ties), starting from dist/assign as given when `accumulate` != 0, else from +inf / 0.
 * frame_idx (optional, may be NULL) restricts the pass to a subset of local frames -- the
 * X[dst_up_assig_this] re-assignment of kmedoids.py:666-667.  With scatter == 0 out_dist /
 * out_assign have one entry per visited frame, in visiting order (length n_idx, or n when
 * frame_idx is NULL); with scatter != 0 they are full-length arrays and the result of frame f
 * is written at position f (the new_dist[mask] = ... scatter of kmedoids.py:669-670). */
int eb_rmsd_assign(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                   const float *centers_soa, const double *center_traces, int32_t k,
                   const int64_t *frame_idx, int64_t n_idx, float *out_dist,
                   int32_t *out_assign, int accumulate, int scatter, void *stream);
/* n_idx = host upper bound of the subset size, *n_idx_dev (int32, optional) the real one. */
int eb_rmsd_assign_dev(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                       const float *centers_soa, const double *center_traces, int32_t k,
                       const int64_t *frame_idx, int64_t n_idx, float *out_dist,
                       int32_t *out_assign, int accumulate, int scatter,
                       const int32_t *n_idx_dev, void *stream);

size_t eb_feat_record_bytes(int64_t n_features, int dtype);
/* ---- K2: fused k-centers step, feature vectors (libdist.pyx:100-145) -------------------
 * X is (n, F) row-major of `dtype`; dist is float64 like libdist's output; arithmetic follows
 * the generated C of the reference bit for bit: typed difference, typed square, float64
 * accumulation in j order, double sqrt. */
int eb_kcenters_step_feat(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                          int64_t frame_offset, const void *cand_in, int n_cand, double *dist,
                          int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                          eb_kc_state *state, int64_t *center_list, void *partials,
                          void *cand_out, int n_steps, void *stream);
int eb_kcenters_seed_feat(const void *X, int64_t n, int64_t n_features, int dtype,
                          int64_t frame_offset, const double *dist, int32_t first_center_id,
                          eb_kc_state *state, void *partials, void *cand_out, void *stream);
/* libdist.euclidean / manhattan(X, y) itself: out float64[n] */
int eb_feat_one_to_all(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                       const void *y, double *out, void *stream);

/* md.rmsd(X, proposal) for a PAM proposal (kmedoids.py:637) with triangle-inequality pruning:
 * out[f] = +inf, and frame f is not read, when assign[f] != cid and
 * dist[f] <= (1 - 1e-5)/2 * cc[assign[f]] with cc[j] = d(proposal, medoid j); every other
 * entry equals eb_rmsd_one_to_all's.  The split of kmedoids.py:644-658 leaves a pruned frame
 * unchanged, exactly as its true distance would. */
int eb_rmsd_one_to_all_pruned(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              const float *center_soa, const double *center_trace,
                              const float *dist, const int32_t *assign, const float *cc,
                              int32_t cid, float *out, void *stream);

/* ---- K3 on the tensor cores: split-FP16 tcgen05 screen + exact re-score ------------------
 * Same result as eb_rmsd_assign (dense pass over all n frames, no accumulate): a tcgen05 GEMM
 * (operands x*2^8 = h1 + h2 in FP16, D = A1.B1 + A1.B2 + A2.B1 accumulated in FP32 in TMEM)
 * with a fused QCP epilogue bounds every (frame, centre) distance; only centres whose lower
 * bound does not exceed the frame's best upper bound survive (error model: |d msd| <=
 * kappa * sqrt(Ga*Gb) / n_atoms) and are re-scored exactly in float64 with the reference's
 * strict-'<' centre order.  cand_count[f] receives the number of centres re-scored for frame f,
 * or -1 if a candidate list overflowed (32 per list); those frames are NOT written and must be
 * sent through eb_rmsd_assign (frame_idx subset).  frame_idx (optional, int64[n]) restricts the
 * pass to frames frame_idx[0..n) of xyz_soa (PAM's X[dst_up_assig_this], kmedoids.py:666-667):
 * cand_count is indexed by position, results go to position i or, with `scatter`, to
 * frame_idx[i].  scratch: eb_tc_scratch_bytes(n, n_atoms, k).  Any n_atoms (the packed
 * operand images are zero-padded to a multiple of 32 atoms).  mode 0 is a debugging aid: dbg (n*k*9 floats) receives the
 * approximate inner-product matrices and nothing else is written; mode 2 is a timing probe
 * (pack + screen with the QCP epilogue removed; no outputs). */
size_t eb_tc_scratch_bytes(int64_t n, int n_atoms, int32_t k);
int eb_rmsd_assign_tc(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                      const float *centers_soa, const double *center_traces, int32_t k,
                      double kappa, const int64_t *frame_idx, int scatter, float *out_dist,
                      int32_t *out_assign, int32_t *cand_count, void *scratch, float *dbg,
                      int mode, void *stream);
/* The same pass when the number of frames is known only on the DEVICE (PAM: the ambiguous
 * subset is counted by eb_pam_classify): n is the host's upper bound (grids, scratch and list
 * layout are sized by it), *n_dev (int32, <= n, optional) the real count -- positions beyond it
 * are neither read nor written.  overflow_count (optional) is incremented once per frame whose
 * candidate lists overflowed, so the caller can defer the exact fallback to its next read-back
 * instead of synchronising after every pass. */
int eb_rmsd_assign_tc_dev(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          const float *centers_soa, const double *center_traces, int32_t k,
                          double kappa, const int64_t *frame_idx, int scatter, float *out_dist,
                          int32_t *out_assign, int32_t *cand_count, void *scratch, float *dbg,
                          int mode, const int32_t *n_dev, int32_t *overflow_count, void *stream);
/* Exact scoring of explicit per-position centre lists (the re-score stage on its own; exact
 * float64 inner products, lowest centre index on ties like cluster/util.py:201): position p
 * covers frame frame_idx[p] (or p when NULL) and the centres cand_list[p*list_len ..
 * +cand_count[p]); out_dist / out_assign[p] receive the nearest listed centre.  The host uses it
 * to AUDIT the tensor-core screen: a sample of frames is scored against every centre and must
 * reproduce what eb_rmsd_assign_tc wrote.  bound_lo: n_pos*list_len floats of -inf, bound_up:
 * n_pos floats of +inf, zero_flag: one int32 holding 0 (all device memory). */
int eb_rmsd_score_lists(const float *xyz_soa, const double *traces, int64_t n_pos, int n_atoms,
                        const float *centers_soa, const double *center_traces,
                        const int32_t *cand_count, const int32_t *cand_list, int list_len,
                        const float *bound_lo, const float *bound_up, const int32_t *zero_flag,
                        const int64_t *frame_idx, float *out_dist, int32_t *out_assign,
                        int32_t *frame_flag, void *stream);

/* ---- K3 for feature vectors (same contract as eb_rmsd_assign) */
int eb_feat_assign(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                   const void *centers, int32_t k, const int64_t *frame_idx, int64_t n_idx,
                   double *out_dist, int32_t *out_assign, int accumulate, int scatter,
                   void *stream);

/* ---- K4/K6: PAM bookkeeping (cluster/kmedoids.py:611-694) -------------------------------
 * eb_pam_classify: given the full-pass distances to the proposal (new_ctr_dist) builds
 *   new_dist/new_assign for the 'dn' and 'up_other' cases, compacts the 'up_this' frame
 *   indices (ascending) into ambig_idx and counts them in *n_ambig.
 * eb_pam_cost:     sum(dist^2) in float64, deterministic order (kmedoids.py:478-479 numerator).
 * eb_count_members / eb_select_member: |{i: assign[i]==cid}| and the kth such i (ascending),
 *   i.e. np.where(assignments == cid)[0][kth] of kmedoids.py:611 + :514.
 * DT = 0 for float32 distance arrays (RMSD), 1 for float64 (feature metrics). */
int eb_pam_classify(const void *new_ctr_dist, const void *dist, const int32_t *assign, int64_t n,
                    int dist_is_f64, int32_t cid, void *new_dist, int32_t *new_assign,
                    int64_t *ambig_idx, int64_t *n_ambig, void *stream);
/* Triangle-inequality pre-pass of a proposal's full distance pass (kmedoids.py:637; RMSD is a
 * metric): with cc[j] = d(proposal, medoid j), a frame of another cluster with
 * dist <= (1 - 1e-5)/2 * cc[assign] provably keeps its medoid -- out[f] = +inf, never read --
 * every other frame index goes to need_idx (compact, unordered), *n_need receives their number.
 * The caller evaluates the proposal against exactly those frames (eb_rmsd_assign_dev, k = 1,
 * scattered output), so the split of kmedoids.py:644-658 sees the same values as after a full
 * pass. */
int eb_pam_need_list(const float *dist, const int32_t *assign, const float *cc, int64_t n,
                     int32_t cid, float *out, int64_t *need_idx, int64_t *n_need, void *stream);
/* scratch of eb_sum_squares / eb_select_member: eb_pam_scratch_bytes(n) bytes, ZERO-INITIALISED
 * once by the caller (eb_sum_squares keeps a ticket counter behind its partial sums and leaves
 * it zero) */
size_t eb_pam_scratch_bytes(int64_t n);
int eb_sum_squares(const void *dist, int64_t n, int dist_is_f64, double *out, void *scratch,
                   void *stream);
int eb_count_members(const int32_t *assign, int64_t n, int32_t k, int64_t *counts, void *stream);
int eb_select_member(const int32_t *assign, int64_t n, int32_t cid, int64_t kth, int64_t *out,
                     void *scratch, void *stream);

/* One PAM proposal for an RMSD shard, queued with ONE call (the sequence the host layer
 * otherwise issues as ~25 separate launches / copies; kmedoids.py:609-694 for one cluster id).
 * All pointers are device memory owned by the caller except pin_*, which are pinned host
 * memory.  `stages` is a bit mask:
 *   EB_PAM_SELECT  the kth member of cluster cid (eb_select_member) is gathered into the
 *                  proposal slot and prop_idx[0] = its GLOBAL index (the owner rank's part,
 *                  kmedoids.py:482-517; a sharded run broadcasts the slot afterwards);
 *   EB_PAM_TRIAL   cc = d(proposal, every medoid); triangle-inequality need list + exact
 *                  distances proposal -> those frames; three-way split (eb_pam_classify);
 *                  scal_i = {prop_idx, n_ambig}; medoid cid saved and replaced by the proposal;
 *                  the ambiguous frames (at most m_max, real count on the device) re-assigned
 *                  against all k medoids (tensor-core screen + exact re-score when use_tc, else
 *                  the exact kernel) -- or, with use_list, against the medoids the triangle
 *                  inequality leaves: medoid j can only win frame x if d(proposal, m_j) <
 *                  2 d(x, proposal), and both numbers are already known (cc and new_ctr_dist),
 *                  so the listed medoids (a few per cent of k on clustered data) are scored
 *                  exactly and everything else is skipped; same result.  More than
 *                  med_list_cap listed medoids count as an overflow in tc_ovf and the caller
 *                  re-assigns the subset itself;
 *                  scal_d = sum(new_dist^2);
 *   EB_PAM_READBACK scal_d, scal_i and tc_ovf copied to pin_d / pin_i / pin_o (asynchronous;
 *                  the caller synchronises the stream once).
 * eb_pam_restore_medoid puts the saved medoid back after a rejected proposal. */
#define EB_PAM_SELECT 1
#define EB_PAM_TRIAL 2
#define EB_PAM_READBACK 4
typedef struct eb_pam_ctx {
    const float *xyz;            /* shard, SoA (n, 3, A_pad) */
    const double *traces;
    int64_t n;
    int64_t frame_offset;        /* global index of local frame 0 */
    int32_t n_atoms;
    int32_t k;
    float *medoid_xyz;           /* (k, 3, A_pad) */
    double *medoid_traces;
    float *prop_xyz;             /* proposal slot, one frame */
    double *prop_traces;
    int64_t *prop_idx;           /* its global index */
    float *saved_xyz;            /* the medoid the proposal displaces */
    double *saved_traces;
    const float *dist;           /* current state */
    const int32_t *assign;
    float *new_dist;             /* trial state */
    int32_t *new_assign;
    float *new_ctr_dist;         /* (n) distances to the proposal, +inf where pruned */
    float *cc;                   /* (k) */
    int64_t *need_idx;           /* (n) */
    int64_t *need_n;             /* (1) */
    int32_t *need_assign;        /* (n) */
    int64_t *ambig_idx;          /* (n) */
    int64_t *scal_i;             /* (2) {proposal, n_ambig} */
    double *scal_d;              /* (1) */
    void *scratch;               /* eb_pam_scratch_bytes */
    int32_t *tc_cand;            /* tensor-core screen workspace for m_max frames */
    void *tc_scratch;
    int32_t *tc_ovf;             /* (1) frames whose candidate lists overflowed */
    double kappa;
    int32_t use_tc;
    int32_t reserved;
    double *pin_d;               /* pinned host: (1), (2), (1) */
    int64_t *pin_i;
    int32_t *pin_o;
    int32_t *med_list;           /* (med_list_cap) medoids that can still win an ambiguous frame */
    int32_t *med_list_n;         /* (1) */
    int32_t med_list_cap;
    int32_t use_list;            /* re-assign the ambiguous frames against med_list only */
} eb_pam_ctx;
int eb_pam_propose_rmsd(const eb_pam_ctx *ctx, int32_t cid, int64_t kth, int64_t m_max,
                        int stages, void *stream);
int eb_pam_restore_medoid(const eb_pam_ctx *ctx, int32_t cid, void *stream);
/* sizeof(eb_kc_state) / sizeof(eb_pam_ctx) as compiled: a binding checks its struct mirrors */
size_t eb_struct_bytes(int which /* 0: eb_kc_state, 1: eb_pam_ctx */);

/* ---- trajectory input without mdtraj: native GROMACS .xtc reader (host code) -------------
 * Replaces md.load(path, stride=, atom_indices=) for .xtc files in the loaders either side of
 * the hot path (cluster/util.py:350-404 load_trajectories, mpi/io.py:142-194
 * load_trajectory_as_striped, the batch loader of cluster/util.py:584-649).  eb_xtc_scan walks
 * the frame headers only.  eb_xtc_read decodes frames first, first+stride, ... (at most
 * max_frames; < 0: all) into xyz_out[(frame, atom, 3)] float32 nanometres (HOST memory);
 * atom_idx (n_sel indices, optional) selects and orders the atoms while decoding. */
int eb_xtc_scan(const char *path, int64_t *n_frames, int32_t *n_atoms);
int eb_xtc_read(const char *path, int64_t first, int64_t stride, int64_t max_frames,
                const int32_t *atom_idx, int32_t n_sel, float *xyz_out, int64_t *n_read);

/* ---- synthetic data: counter-based generator keyed on (seed, global frame) -
 * Writes centred SoA frames + traces directly in HBM so a 10M x 500-atom trajectory never has
 * to exist on the host.  The same generator is restated in numpy (enspara_b200/synth.py) for
 * down-scaled oracle comparisons; eb_synth_trajectory_aos writes the raw (n,A,3) frames. */
int eb_synth_trajectory_aos(float *xyz_aos, int64_t n, int n_atoms, int64_t first_frame,
                            uint64_t seed, const float *base_conformers, int n_base,
                            void *stream);
int eb_synth_features(float *X, int64_t n, int64_t n_features, int64_t first_row, uint64_t seed,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ENSPARA_B200_H */

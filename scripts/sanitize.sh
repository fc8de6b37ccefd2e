#!/bin/bash
# compute-sanitizer passes over the hot path: memcheck / racecheck /
# synccheck on one small invocation of every kernel family (smoke(): RMSD k-centers + PAM +
# euclidean k-centers) and on the emulated multi-rank peer-memory exchange.  Writes
# gpurun_out/sanitize_<tool>_<what>.log; the ERROR SUMMARY lines are what profiles/ keeps.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
export EB_EXCH_TIMEOUT_S=600
for tool in memcheck racecheck synccheck; do
  timeout 240 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" \
      > "$OUT/sanitize_${tool}_smoke.log" 2>&1
  echo "$tool smoke rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_${tool}_smoke.log" | tail -1)"
  timeout 240 $CS --tool $tool --print-limit 20 python -m pytest -q -x \
      tests/test_gpu_exchange_single.py::test_overlapping_ranks_small_shards \
      tests/test_gpu_exchange_single.py::test_empty_shards_and_stop_rule \
      > "$OUT/sanitize_${tool}_exchange.log" 2>&1
  echo "$tool exchange rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_${tool}_exchange.log" | tail -1)"
done
# tensor-core screen + re-score (tcgen05 / TMEM / mbarrier ring) and the TMA step kernel
timeout 240 $CS --tool memcheck --print-limit 20 python -m pytest -q -x \
    "tests/test_gpu_tc_screen.py::test_tc_assign_equals_exact_assign" \
    > "$OUT/sanitize_memcheck_tc.log" 2>&1
echo "memcheck tc rc=$? : $(grep -E 'ERROR SUMMARY' "$OUT/sanitize_memcheck_tc.log" | tail -1)"
timeout 240 $CS --tool racecheck --print-limit 20 python -m pytest -q -x \
    "tests/test_gpu_tc_screen.py::test_tc_assign_equals_exact_assign" \
    > "$OUT/sanitize_racecheck_tc.log" 2>&1
echo "racecheck tc rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_racecheck_tc.log" | tail -1)"
# round-2 additions: the persistent multi-iteration kernels (grid barrier, dynamic tail, TMA ring
# across iterations), the compact pruned pass of PAM, the re-score bucketing, the xtc-free CLI
for tool in memcheck racecheck; do
  timeout 400 $CS --tool $tool --print-limit 20 python -m pytest -q -x \
      "tests/test_gpu_features.py::test_multi_and_single_launches_mix" \
      "tests/test_gpu_features.py::test_multi_iteration_kernel_edge_cases" \
      "tests/test_gpu_kcenters_rmsd.py::test_persistent_multi_iteration_kernel_equals_single_launches[3001-50]" \
      "tests/test_gpu_kmedoids.py::test_pam_pruned_full_pass_changes_nothing" \
      "tests/test_gpu_tc_screen.py::test_tc_audit_every_call_and_detects_corruption" \
      > "$OUT/sanitize_${tool}_round2.log" 2>&1
  echo "$tool round2 rc=$? : $(grep -E 'passed|failed' "$OUT/sanitize_${tool}_round2.log" | tail -1) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_${tool}_round2.log" | tail -1)"
done
timeout 400 $CS --tool memcheck --print-limit 20 python -m pytest -q -x \
    "tests/test_gpu_tc_screen.py::test_tc_assign_matches_oracle_other_atom_counts" \
    > "$OUT/sanitize_memcheck_bucketing.log" 2>&1
echo "memcheck bucketing rc=$? : $(grep -E 'passed|failed' "$OUT/sanitize_memcheck_bucketing.log" | tail -1) $(grep -E 'ERROR SUMMARY' "$OUT/sanitize_memcheck_bucketing.log" | tail -1)"

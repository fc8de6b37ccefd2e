#!/bin/bash
# Round-2 single-GPU measurement campaign (run under gpurun): tests, bench (both arms), every
# BASELINE config through the public API with the CPU baselines beside it, ncu launch list and
# full captures of the dominant kernels.  Everything lands in gpurun_out/ (r2_* names).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_gputests.log; cat $O/r2_gputests.log
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 600 $O/r2_bench_n1.json; echo
python bench.py --impl reference --steps 20 --warmup 3 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -c 400 $O/r2_bench_reference.json; echo
python scripts/bench_configs.py --only peaks,c1,c2,c3,c5,tri,reassign --out $O/r2_configs.json > $O/r2_configs.log 2>&1; tail -c 300 $O/r2_configs.log; echo
nvidia-smi topo -m > $O/r2_topo.txt 2>&1
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_launches_bench.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:k_kcenters_step_rmsd_tma -s 5 -c 1 \
    -o $O/r2_k1_prof python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_k1_prof.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:"k_tc_screen|k_rescore" -s 2 -c 2 \
    -o $O/r2_tc_prof python scripts/dev_prof_tc2.py > $O/r2_tc_prof.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:k_kcenters_multi_feat -c 1 \
    -o $O/r2_k2_prof python scripts/dev_prof_k2.py > $O/r2_k2_prof.log 2>&1
ls -la $O/*.ncu-rep

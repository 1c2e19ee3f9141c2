#!/bin/bash
# One gpurun call for the round's records: GPU tests, smoke, bench (ours, reference arm, sweep / chamfer / action
# workloads), ncu launch list of the bench command and ncu --set full captures of one launch per hot kernel.
#   gpurun --timeout 1500 -- 'bash tools/final_check.sh r02k'
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --workload sweep --steps 5 --warmup 3 > $OUT/bench_sweep_$TAG.json 2> $OUT/bench_sweep_$TAG.err; echo "sweep rc=$?"
python bench.py --workload chamfer --steps 5 --warmup 3 > $OUT/bench_chamfer_$TAG.json 2> $OUT/bench_chamfer_$TAG.err; echo "chamfer rc=$?"
python bench.py --workload action --steps 10 --warmup 3 > $OUT/bench_action_$TAG.json 2> $OUT/bench_action_$TAG.err; echo "action rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
if [ "${SKIP_FULL:-0}" != "1" ]; then
python tools/prof_kernels.py > $OUT/prof_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"knn_feat|feat_split|group_|fps_reg|ball_query_kernel|grid_knn|grid_nn1|csr_cluster|edge_affine" -c 80 -f -o $OUT/prof_$TAG python tools/prof_kernels.py > $OUT/ncu_prof_$TAG.log 2>&1
echo "ncu full rc=$?"
fi

#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (ours + reference arm), ncu launch list and
# full captures of the hot kernels.  Outputs land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
python bench.py --steps 10 --warmup 3 --per-op > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.err | tail -15; cat $OUT/bench_$TAG.json
if [ "${SKIP_REF:-0}" != "1" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
for K in ${NCU_KERNELS:-knn_warp_kernel group_fwd_kernel fps_reg_kernel nn1_kernel group_bwd_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-3} -c 2 -f -o $OUT/prof_${K}_$TAG $CMD > $OUT/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K rc=$?"
done
fi

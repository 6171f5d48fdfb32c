#!/bin/bash
# One gpurun call that regenerates every artefact profiles/README.md quotes (run from the repo root on a B200 box):
#   bash tools/capture_profiles.sh r2
# writes gpurun_out/*_${TAG}*; copy / condense into profiles/ with tools/ncu_export.py and tools/ncu_traffic.py.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_${TAG}_final.json 2> $OUT/bench_${TAG}_final.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-train \
    --no-sliding-window --no-cpu-baseline --no-parity --preheat 0 > $OUT/launches_${TAG}.log 2>&1
for what in attn attn0 gemm_big gemm_gelu gemm_res pool_qkv0 pool_qkv4; do
  timeout 300 ncu --set full --import-source on --clock-control none -s 2 -c 1 -k regex:'attention_tc|linear_tc|pool_tma' \
      -f -o $OUT/prof_${TAG}_${what} python tools/profile_one.py $what > /dev/null 2>&1
done
python tools/determinism_probe.py --iters 100 --out $OUT/determinism_${TAG}.json > $OUT/determinism_${TAG}.txt 2>&1
python tools/attn_determinism.py --runs 10 --out $OUT/attn_determinism_${TAG}.json > $OUT/attn_determinism_${TAG}.txt 2>&1
python tools/microbench.py --json $OUT/micro_${TAG}_final.json > $OUT/micro_${TAG}_final.txt 2>&1
python tools/attn_bench.py > $OUT/attn_bench_${TAG}.txt 2>&1
python tools/lnfold_bench.py $OUT/lnfold_${TAG}.json > $OUT/lnfold_${TAG}.txt 2>&1
ls -la $OUT | tail -20

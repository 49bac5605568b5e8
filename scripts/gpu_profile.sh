#!/bin/bash
# One gpurun call (1 GPU) that refreshes the ncu evidence under gpurun_out/ (summaries: profiles/summarize.py):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh r2'
# Every profiled command line first runs without ncu in the same call (the `&&` before ncu).
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
small="--steps 2 --warmup 3 --no-extra --no-cpu-baseline --latency-queries 3 --no-parity"
one="--steps 1 --warmup 3 --no-extra --no-cpu-baseline --latency-queries 3 --no-parity"
# launch list of the headline step (every kernel with its device time; cold-cache, serialised: compare shares)
python bench.py $small > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py $small > $out/ncu_launches_$tag.log 2>&1
echo "launch list exit=$?"
# headline: the long-launch CTA-pair instantiation gemm_topk_kernel<1, 1, 0> (4th launch)
python bench.py $one > $out/plain1_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel --launch-skip 3 -c 1 \
    -o $out/prof_${tag}_gemm_long -f python bench.py $one > $out/ncu_gemm_long_$tag.log 2>&1
echo "ncu gemm long exit=$?"
# the N = 8 shard: the short-launch instantiation gemm_topk_kernel<1, 1, 1>
python bench.py $one --rows 1250000 > $out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel --launch-skip 3 -c 1 \
    -o $out/prof_${tag}_gemm_short -f python bench.py $one --rows 1250000 > $out/ncu_gemm_short_$tag.log 2>&1
echo "ncu gemm short exit=$?"
# the tail of a step: selection kernel and the fused rerank tail (4th launch of each)
ncu --set full --clock-control none --import-source on -k regex:select_warp_kernel --launch-skip 3 -c 1 \
    -o $out/prof_${tag}_select -f python bench.py $one > $out/ncu_select_$tag.log 2>&1
echo "ncu select exit=$?"
ncu --set full --clock-control none --import-source on -k regex:rerank_scored_kernel --launch-skip 3 -c 1 \
    -o $out/prof_${tag}_rerank -f python bench.py $one > $out/ncu_rerank_$tag.log 2>&1
echo "ncu rerank exit=$?"
# batch-1 scan
ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel --launch-skip 4 -c 1 \
    -o $out/prof_${tag}_scan -f python bench.py $one > $out/ncu_scan_$tag.log 2>&1
echo "ncu scan exit=$?"
ls -la $out/*.ncu-rep

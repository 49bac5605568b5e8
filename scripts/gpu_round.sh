#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU tests, the default bench line, the ncu
# launch list of the same command and one `--set full` capture of each dominant kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r1'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit=$?" >> $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit=$?"
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit=$?"
cat $out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
cat $out/bench_ref_$tag.json
python scripts/eval_cfg5.py > $out/cfg5_$tag.json 2>$out/cfg5_$tag.err; cat $out/cfg5_$tag.json
python bench.py --rows 1000000 --batch 1024 --no-cpu-baseline --latency-queries 0 > $out/bench_cfg2_$tag.json 2>>$out/bench_$tag.err; cat $out/bench_cfg2_$tag.json
small="--steps 2 --warmup 3 --no-cpu-baseline --latency-queries 3"
python bench.py $small > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py $small > $out/ncu_launches_$tag.log 2>&1
echo "launch list exit=$?"
one="--steps 1 --warmup 3 --no-cpu-baseline --latency-queries 3"
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel --launch-skip 3 -c 1 \
    -o $out/prof_${tag}_gemm -f python bench.py $one > $out/ncu_gemm_$tag.log 2>&1
echo "ncu gemm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel --launch-skip 4 -c 1 \
    -o $out/prof_${tag}_scan -f python bench.py $one > $out/ncu_scan_$tag.log 2>&1
echo "ncu scan exit=$?"

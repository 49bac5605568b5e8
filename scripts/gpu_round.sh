#!/bin/bash
# One gpurun call (1 GPU) that refreshes a round's evidence: GPU tests, smoke, the default bench line, the
# reference arm, then the ncu launch list and `--set full` captures (scripts/gpu_profile.sh).
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit=$?" >> $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit=$?"
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
cat $out/bench_ref_$tag.json
bash scripts/gpu_profile.sh $tag

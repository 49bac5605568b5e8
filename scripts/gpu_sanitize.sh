#!/bin/bash
# compute-sanitizer over the small GPU test matrix (ONE tool per gpurun call, see /opt/skills/guides/B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh memcheck'
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh racecheck'
tool=${1:-memcheck}
out=gpurun_out; mkdir -p $out
sel='test_edge_shapes or test_gemm_other_dims or test_gemm_cta_pairs or test_exclude_rows_and_row_offset or test_gemm_duplicates_and_zero_rows or test_fused_tail_vs_oracle or test_exchange_kernels_on_one_gpu or test_metrics_match_reference_golden or test_rerank_matches_reference_golden or test_merge_topk_matches_single_shard or test_diversity or test_label_relevance'
if [ "$tool" = racecheck ]; then
  # shared-memory hazards: the kernels with hand-rolled smem protocols on their smallest cases
  sel='test_edge_shapes or test_gemm_other_dims or test_fused_tail_vs_oracle or test_exchange_kernels_on_one_gpu or test_merge_topk_matches_single_shard'
fi
# the plain run first (a faulting program must not be put under the tool)
python -m pytest tests/test_gpu_search.py tests/test_gpu_rerank_metrics.py -m gpu -q -x -k "$sel" > $out/sanitize_plain_$tool.log 2>&1 || { tail -5 $out/sanitize_plain_$tool.log; echo "plain run failed"; exit 1; }
tail -1 $out/sanitize_plain_$tool.log
timeout 1300 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
    python -m pytest tests/test_gpu_search.py tests/test_gpu_rerank_metrics.py -m gpu -q -x -k "$sel" > $out/sanitize_$tool.log 2>&1
echo "compute-sanitizer $tool exit=$?"
grep -E "ERROR SUMMARY|passed|failed|Error|RACECHECK SUMMARY" $out/sanitize_$tool.log | tail -8

#!/bin/bash
# compute-sanitizer (one tool per call) on small cases of every kernel
mkdir -p gpurun_out
TOOL=${1:-memcheck}
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 3 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x \
  -k "dmma_gemm_matches_plain_fp64 and not 2048 or spd_factor_and_inverse and (127 or 129 or 300) or nll_grad_matches_oracle and 64 or fill_edge or literal_inputs or applymap_matches_oracle and pendulum or nan_and_empty or split_map or explicit_map_matches_reference" \
  > gpurun_out/sanitize_$TOOL.log 2>&1
echo "sanitizer $TOOL exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error|error" gpurun_out/sanitize_$TOOL.log | head -20
tail -5 gpurun_out/sanitize_$TOOL.log

#!/usr/bin/env bash
# compute-sanitizer passes over the hand-written kernels (race detection).  Run on a B200 box:
#   bash tools/sanitize.sh        -> gpurun_out/sanitize_{memcheck,racecheck,synccheck}.txt (summaries copied to profiles/)
# memcheck: out-of-bounds / misaligned global + shared accesses; racecheck: shared-memory hazards between threads;
# synccheck: invalid barrier usage.  The test subset touches every kernel family once at small shapes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='not 2560 and not 2500 and not 8192 and (test_attn_sparse3dna_core or test_attn_dense_x64_two_pass_kernel or test_gemm_matches or test_gemm_epilogues or (umma and 601) or (halo and 601) or test_vq_argmax_tensor_core_path_is_the_fp32_argmax or test_sandwich_ln_shift_scatter or test_fused_decode_model_geometry)'
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py \
    tests/test_decode_kernels_gpu.py tests/test_vae_gpu.py tests/test_fused_decode_gpu.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_$tool.txt | tail -5
done

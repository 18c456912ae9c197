#!/usr/bin/env bash
# First GPU call of the next session: run every opt-in (never-run-on-hardware) kernel's test on ONE GPU, each under
# its own timeout so a hung mbarrier costs two minutes, not the box.  Order = cheapest / most fundamental first.
#   gpurun --timeout 900 -- 'bash scripts/validate_experimental.sh'
cd "$(dirname "$0")/.."
run() {  # name, env assignment, pytest selector
  echo "=== $1"
  env $2 timeout 150 python -m pytest "$3" -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | grep -vi warning | tail -6
  echo "    rc=${PIPESTATUS[0]}"
}
run "TMA im2col probe (semantics of cuTensorMapEncodeIm2col coordinates)" BATON_TMA_IM2COL=1 \
    "tests/test_gpu_kernels.py::test_tma_im2col_probe_matches_explicit_im2col"
run "implicit-GEMM convolution forward + wgrad" BATON_CONV_IGEMM=1 \
    "tests/test_gpu_kernels.py::test_implicit_gemm_conv_matches_im2col_path"
run "single-kernel BatchNorm backward (device-wide generation barrier)" BATON_BN_BWD_FUSED=1 \
    "tests/test_gpu_kernels.py::test_batchnorm_backward_single_kernel_matches_two_kernel_path"
run "fused attention forward + backward (S=128, d=64)" BATON_FUSED_ATTN=1 \
    "tests/test_gpu_bert.py::test_fused_attention_forward_and_backward_match_multi_kernel_path"
echo "=== bench with everything that passed switched on (edit the env list to what passed)"
echo "BATON_BN_BWD_FUSED=1 BATON_CONV_IGEMM=1 python bench.py --steps 10 --warmup 3"
echo "BATON_FUSED_ATTN=1 python bench.py --model bert_base --steps 3 --warmup 3"

#!/bin/bash
# compute-sanitizer passes over the warp-specialised tcgen05 / TMA / mbarrier kernels at CI-size shapes.
# usage (GPU box): tools/sanitize.sh <out-dir>        -> <out-dir>/sanitizer_{memcheck,racecheck,synccheck}.log
# The selected tests cover tc_gemm (LinearNT / LinearTN / Conv2HeadsTC / GenL1*), tc_gemm2 (Conv1Fwd / Conv1Wgrad / GenL1*Pair with
# Fourier features and with the fused coordinate layer / LinearTNPair: store-issuer warp, staged / freed mbarriers, TMA
# reduce-add epilogues), enc_heads_bwd / enc_dx1_dw2 (enc_bwd_fused.cuh), the attention cluster kernel and the refinement
# kernels; the golden / edge-shape / resid steps also run the streaming kernels (thin_bwd_stream, group_colsum8, the filter-bank
# plane / gather kernels) and the one-bit mask writers (GenL1FwdPair and LinearNT epilogues); one CUDA-graph capture + replay.
out=${1:-gpurun_out}
mkdir -p "$out"
SEL='test_linear_nt[128-128-64] or test_linear_tn[64-128-128-0] or test_linear_nt_full_epilogue[dgrad-300-256-256] or test_linear_tn_pair_matches_tc_gemm[999-512-256] or test_step_matches_reference_golden[g1_mnist] or test_step_matches_reference_golden[g4_particles_ctf] or test_get_latent_matches_reference_golden[g1_mnist] or test_generator_coord_fused_matches_unfused[2-False] or test_resid_generator_module_matches_oracle or test_edge_shapes_match_oracle or test_replay_equals_eager[cfg4_graph]'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_gemm_core.py tests/test_gpu_step.py tests/test_gpu_argmax.py tests/test_gpu_stages.py tests/test_gpu_graph.py -q -x -k "$SEL" > "$out/sanitizer_$tool.log" 2>&1
  echo "exit code $?" >> "$out/sanitizer_$tool.log"
  grep -E "ERROR SUMMARY|passed|failed|exit code" "$out/sanitizer_$tool.log" | tail -4
done

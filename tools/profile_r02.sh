#!/bin/bash
# Final-revision ncu captures (run under gpurun on ONE GPU).  Each target first runs plain (must exit 0), then under
# `ncu --set full --clock-control none` for ONE launch of the named kernel; reports land in gpurun_out/r02_<target>.ncu-rep.
set -u
mkdir -p gpurun_out
prof() {   # target kernel-regex skip
  python tools/profile_targets.py "$1" > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o gpurun_out/r02_$1 \
      python tools/profile_targets.py "$1" > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
for t in "$@"; do
  case $t in
    encoder_L8)  prof encoder_L8 encoder_fused 2 ;;
    encoder_L64) prof encoder_L64 encoder_fused 2 ;;
    attn_L8)     prof attn_L8 attn_tc5 2 ;;
    attn_L64)    prof attn_L64 attn_tc5 2 ;;
    attn_L256)   prof attn_L256 attn_tc5 2 ;;
    conv_tap)    prof conv_tap gemm_bf16_tn 8 ;;
    ln_film)     prof ln_film ln_film 2 ;;
    embed)       prof embed embed_traj 1 ;;
    sgemm)       prof sgemm sgemm 2 ;;
    interp_T256) prof interp_T256 nested_masks 2 ;;
    interp_T64)  prof interp_T64 nested_masks 2 ;;
    gemm_qkv384) prof gemm_qkv384 gemm_bf16_tn 2 ;;
    corrupt_adj) prof corrupt_adj corrupt_adjacent 2 ;;
    ln_bwd)      prof ln_bwd ln_bwd_apply 2 ;;
    qkv_attn)    prof qkv_attn qkv_attn_kernel 2 ;;
    mlp_pair)    prof mlp_pair mlp_pair_kernel 2 ;;
    im2col)      prof im2col im2col3x3_scatter 2 ;;
    colsum)      prof colsum colsum_partial_vec 2 ;;
  esac
done

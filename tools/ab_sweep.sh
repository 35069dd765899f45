#!/bin/bash
# A/B of two builds of the library on ONE box (boxes differ by a few per cent): tools/ab_sweep.sh <other libqrusty_cuda.so>
# Build the other one from another checkout, e.g.
#   git --work-tree=/tmp/prev checkout <rev> -- qrusty_b200/csrc include && (cd /tmp/prev/qrusty_b200/csrc && nvcc <FLAGS of build.py> -o <repo>/qrusty_b200/lib/libqrusty_cuda_prev.so qrusty_cuda.cu -ldl)
# (files under qrusty_b200/lib/ travel to the GPU box with gpurun and stay out of git).
OTHER=${1:-/root/repo/qrusty_b200/lib/libqrusty_cuda_prev.so}
mkdir -p gpurun_out
S=gpurun_out/${TAG:-ab}_sweep.jsonl; : > $S
for rep in 1 2; do
for lib in "" "$OTHER"; do
  export QRUSTY_CUDA_LIB=$lib; [ -z "$lib" ] && unset QRUSTY_CUDA_LIB
  for w in "H8" "H12 --rows 18" "H10 --rows 17" "H11 --rows 16" "C3 --rows 18 --max-gb 10" "rand:22:96:64 --rows 20" "rand:24:6000:3000 --rows 16" "rand:20:1200:600 --rows 17"; do
    timeout 200 python tools/fill_sweep.py $w --reps 20 --cfgs "auto" 2>/dev/null | sed "s|\"cfg\": \"auto\"|\"cfg\": \"auto lib=${lib##*/}\"|" >> $S
  done
done
done
cat $S | cut -c1-75,100-260

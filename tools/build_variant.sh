#!/bin/bash
# tools/build_variant.sh <git-ref> <tag>: compile the library sources of <git-ref> into sr_gan_fd_b200/libb200sr_<tag>.so
# (same-box A/B runs: B200SR_LIB=$PWD/sr_gan_fd_b200/libb200sr_<tag>.so python bench.py ...)
set -e
ref=$1; tag=$2
tmp=$(mktemp -d)
git archive "$ref" sr_gan_fd_b200/csrc include | tar -x -C "$tmp"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -diag-suppress 550 -diag-suppress 177 \
  -o sr_gan_fd_b200/libb200sr_$tag.so "$tmp/sr_gan_fd_b200/csrc/b200sr.cu"
rm -rf "$tmp"
echo built sr_gan_fd_b200/libb200sr_$tag.so

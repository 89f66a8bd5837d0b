#!/bin/bash
# builds tools/_bin/dense_probe_<tag> for each "tag:flags" argument (diagnostic variants of the dense decode kernels)
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -prec-div=true -prec-sqrt=true $flags \
      -o tools/_bin/dense_probe_$tag tools/dense_probe.cu 2> tools/_bin/dense_probe_$tag.log || echo "FAILED $tag" ) &
done
wait
ls -la tools/_bin/

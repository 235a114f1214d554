#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the CPU oracle and (when the reference tree is present) stages the reference's
# own compiled solver under oracle/_ref/.  Nothing in the product path links or loads these.
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p _ref
# C restatement: no FMA contraction, no fast-math, baseline x86-64 (SSE2) like the reference Makefile (-O3, no -march)
gcc -O2 -std=c11 -fPIC -shared -ffp-contract=off -fno-fast-math -fvisibility=hidden \
    -o liblpbox_oracle.so lpbox_oracle.c seg_oracle.c -lm
# the benchmark's workload generator as its own library: bench.py's reference arm must not map the product library
g++ -O2 -std=c++17 -fPIC -shared -o liblpbox_gen.so ../accelerated-lpbox-admm_b200/csrc/auction_gen.cpp -lpthread
REF=/root/reference/Segmentation/Segmentation/cython/src/liblpbox_solver.so
if [ -f "$REF" ]; then
  # The reference sources need Eigen 3.4-dev + OpenCV 4.4 headers (absent here, no network) -> unbuildable.
  # Its shipped x86-64 build of the same sources IS loadable once cv::* is stubbed: stage it unmodified.
  g++ -O1 -fPIC -shared -o _ref/libcvstub.so cvstub.cpp
  if ! cmp -s "$REF" _ref/liblpbox_solver.so 2>/dev/null; then cp -f "$REF" _ref/liblpbox_solver.so; chmod u+w _ref/liblpbox_solver.so; fi
fi
echo "oracle built"

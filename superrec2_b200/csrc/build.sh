#!/bin/bash
# Build libsr2.so in-tree for sm_100a (B200). nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libsr2.so"
nvcc -O3 -std=c++17 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas \
  ${SR2_NVCC_EXTRA:-} \
  -shared -o "$out" "$here"/sr2_context.cu "$here"/sr2_rows.cu "$here"/sr2_dtl.cu "$here"/sr2_superdtl.cu "$here"/sr2_resolve.cu \
  -lgomp -lcudart_static -ldl -lrt -lpthread
echo "built $out"

#!/bin/bash
# compute-sanitizer is closed on the GPU pool ("runs under it have left GPUs needing a reset"), so the
# kernels carry their own invariant checks (SR2_DCHECK, sr2_internal.cuh).  This builds a second copy of
# the library with the checks compiled in and runs the whole GPU test-suite against it; a violated check
# fails its launch (cudaErrorAssert) and the test with it.
#   here:        scripts/debug_checks.sh build
#   GPU box:     scripts/debug_checks.sh run   (through gpurun; writes gpurun_out/debug_checks.log)
set -euo pipefail
root="$(cd "$(dirname "$0")/.." && pwd)"
csrc="$root/superrec2_b200/csrc"
lib="$root/superrec2_b200/libsr2_debug.so"
if [ "${1:-build}" = build ]; then
  nvcc -O2 -std=c++17 -lineinfo -DSR2_DEBUG_CHECKS -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-fopenmp -shared -o "$lib" "$csrc"/sr2_context.cu "$csrc"/sr2_rows.cu "$csrc"/sr2_dtl.cu \
    "$csrc"/sr2_superdtl.cu "$csrc"/sr2_resolve.cu -lgomp -lcudart_static -ldl -lrt -lpthread
  echo "built $lib"
else
  cd "$root"
  SR2_LIBRARY="$lib" python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/debug_checks.log
  SR2_LIBRARY="$lib" python bench.py --workload dtl_batch_small --no-secondary --steps 3 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/debug_checks.log
fi

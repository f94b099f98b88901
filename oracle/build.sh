#!/bin/bash
# Build the CPU oracle (test infrastructure) next to its source.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
gcc -O2 -std=c11 -fopenmp -fPIC -shared -Wall -o "$here/libsr2_oracle.so" "$here/sr2_oracle.c"
echo "built $here/libsr2_oracle.so"

// sr2_superdtl.cu -- the `superdtl` DP (filled in below).
#include "sr2_internal.cuh"

struct SuperProblem {};

void sr2_free_super(sr2_ctx* ctx) {
    delete ctx->super;
    ctx->super = nullptr;
}

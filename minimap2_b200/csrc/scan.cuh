// minimap2_b200/csrc/scan.cuh -- small device prefix-sum helpers (plumbing between the stage kernels).
#pragma once
#include "mmb_internal.h"

// In-place exclusive scan of d[0..n) (int64). If with_total, d[n] receives the total. Returns the total (this
// synchronises the stream: callers use the total to size the next stage's buffers).
int64_t mmb_exclusive_scan_i64(mmb_ctx_t *ctx, int64_t *d, int64_t n, bool with_total);
// Same without returning/synchronising (total written to d[n]).
void mmb_exclusive_scan_i64_async(mmb_ctx_t *ctx, int64_t *d, int64_t n);

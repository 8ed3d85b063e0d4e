// particles.cuh — device-side particle machinery: scans, stable radix binning, cell sort, P2G/G2P/advect.
#pragma once
#include "common.cuh"

// Exclusive prefix sum of n uint32 values (in place allowed: out may equal in). total (optional, device) = sum.
int scan_exclusive_u32(jetcu_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n);

// Stable LSD radix sort of (key, value) pairs; keys < 2^key_bits. Result in keys_out/vals_out.
int radix_sort_pairs_u32(jetcu_ctx* ctx, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                         uint32_t* vals_out, size_t n, int key_bits);

// Entry points shared by capi.cu and group.cu that are NOT part of the public C ABI (include/gcp_b200.h): the chunked
// folds with their result left in DEVICE memory of the context's device (d_out: n_fields x 128 bytes, d_status: n_fields
// bytes), written and stream-synchronised before the call returns.  The host-buffer arguments are the public ones.
#pragma once
#include "../../include/gcp_b200.h"

extern "C" {
int gcp_internal_tally_to_dev(gcp_ctx* ctx, const void* ct, size_t n_ballots, int n_fields, void* d_out, uint8_t* d_status,
                              int fmt);
int gcp_internal_encrypt_tally_to_dev(gcp_ctx* ctx, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                                      int n_fields, void* d_out, uint8_t* d_status, int fmt);
int gcp_internal_ballot_batch_to_dev(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* roots, int shared_root,
                                     const void* siblings, const uint8_t* packed, const uint64_t* offsets, const void* keys,
                                     const void* values, const void* pub_key, const void* k, const void* m, int n_fields,
                                     uint8_t* out_flags, uint8_t* out_status, void* d_tally, uint8_t* d_tally_status, int fmt);
// d_status[f] <- first non-zero of (d_status[f], d_part_status[0..n_parts)[f]); a field with a status gets a zero ciphertext
int gcp_internal_merge_status_dev(gcp_ctx* ctx, const uint8_t* d_part_status, int n_parts, int n_fields, void* d_ct,
                                  uint8_t* d_status, void* stream);
}

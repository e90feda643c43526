// Internal launch interface between the C-ABI host layer (capi.cu) and the device translation unit (kernels.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace gcp {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// per-item status codes (mirrored in include/gcp_b200.h)
enum : u8 {
  GCP_STATUS_OK = 0,
  GCP_STATUS_NONCANONICAL = 1,  // an input element >= r
  GCP_STATUS_KEY_RANGE = 2,     // SMT key >= 2^n_levels (tree/smt/utils.go:11-13)
  GCP_STATUS_NOT_BOOLEAN = 3,   // enabled / fnc / isOld0 not in {0,1}
  GCP_STATUS_OFF_CURVE = 4,     // public key fails AssertIsOnCurve (elgamal/encrypt.go:49)
  GCP_STATUS_ZERO_DENOM = 5,    // Edwards addition denominator is 0 (only reachable off-curve)
  GCP_STATUS_ASSERTION = 6,     // an AssertIsEqual of the gadget fails (SMT processor: old root, LevIns, states, key rule)
  GCP_STATUS_MALFORMED = 7,     // arbo.UnpackSiblings would reject the packed proof (wrapper_arbo.go:64-67)
};

struct PoseidonTable {  // one per t, device pointers into the global-memory copy
  const u32* C;
  const u32* S;
  const u32* M;
  const u32* P;
  const u32* D;  // RP / 2 derived constants of the partial-round pairs (poseidon.cuh), built on the device
  int RP;
  int t;
};

struct SmtArgs {
  int n_levels;
  size_t n;
  const u32* roots;      // n x 8 or 1 x 8 (root_stride = 0)
  size_t root_stride;    // in u32 words: 8 or 0
  const u32* siblings;   // n x n_levels x 8, root -> leaf
  const u32* old_keys;   // n x 8 or nullptr (inclusion form: old == new)
  const u32* old_values; // n x 8 or nullptr
  const u8* is_old0;     // n or nullptr (0)
  const u32* keys;       // n x 8
  const u32* values;     // n x 8 (ignored when fnc = 1: the gadget hashes value but never uses it)
  const u8* fnc;         // n or nullptr (0 = inclusion)
  const u8* enabled;     // n or nullptr (1)
  u32* leaf;             // n x 8 scratch: leaf hash selected by the state machine, lazy Montgomery
  u8* flags;             // n
  u8* status;            // n
  u32* out_roots;        // n x 8 or nullptr: recomputed level[0], canonical
  int mont;              // element format: 0 canonical integers, 1 gnark-crypto Montgomery memory
  int leaf_hash_form;    // VerifierWithLeafHash[Flag] (verifier.go:129-183): `values` / `old_values` hold hash1New / hash1Old
  int hasher;            // utils.Hasher plug: 0 PoseidonHasher, 1 Poseidon2Hasher (utils/hashers.go:25-37)
  const u32* hkeys;      // hasher 1: the context's 62 Poseidon2 round keys (Montgomery form)
};

struct SmtProcessArgs {
  int n_levels;
  size_t n;
  const u32* old_roots;   // n x 8
  const u32* siblings;    // n x n_levels x 8
  const u32* old_keys;    // n x 8
  const u32* old_values;  // n x 8
  const u8* is_old0;      // n
  const u32* new_keys;    // n x 8
  const u32* new_values;  // n x 8
  const u8* fnc0;         // n
  const u8* fnc1;         // n
  u32* new_roots;         // n x 8
  u8* status;             // n
  int mont;
  int leaf_hash_form;     // ProcessorWithLeafHash (processor.go:16): `new_values` / `old_values` hold hash1New / hash1Old
  int hasher;             // as in SmtArgs
  const u32* hkeys;
};

// scratch: perm (n x u32), lidx (n x u16), info (n x u8), hist (256 x u32), cursor (256 x u32)
struct SmtScratch {
  u32* perm;
  u16* lidx;
  u8* info;
  u32* hist;
  u32* cursor;
};

cudaError_t launch_to_mont(u32* d_elems, size_t n, cudaStream_t stream);
cudaError_t upload_const_tables(const u32* d_t3, const u32* d_t4, cudaStream_t stream);
// D[p] = sum_k S_B[k] S_A[t+k-1] for the p-th pair of partial rounds (A = (RP & 1) + 2p, B = A + 1); d_out: RP / 2 elements
cudaError_t launch_poseidon_pair_constants(const PoseidonTable& tab, u32* d_out, cudaStream_t stream);
cudaError_t launch_poseidon(const PoseidonTable& tab, const u32* in, u32* out, u8* status, size_t n_items,
                            int chunks_per_item, size_t in_item_stride, size_t in_chunk_stride,
                            size_t out_item_stride, int in_mont, int out_mont, int final_level,
                            cudaStream_t stream);
double launch_imad_probe(u32* d_out, int blocks, u32 seed, cudaStream_t stream);
cudaError_t launch_smt_verify(const SmtArgs& a, const SmtScratch& sc, int sm_count, cudaStream_t stream);
cudaError_t launch_smt_scan(const u32* siblings, size_t n, int n_levels, u16* lidx, u8* info, u32* hist, int sm_count,
                            cudaStream_t stream);
// sc + acc_old/acc_new (n x 8 words each): the scan / sort / prep / path pipeline; sc == nullptr: one thread per proof
cudaError_t launch_smt_process(const SmtProcessArgs& a, const SmtScratch* sc, u32* acc_old, u32* acc_new, int sm_count,
                               cudaStream_t stream);
// Hash1 rows (tree/smt/hash.go:10-19): rows[i] = (key_i, values_i[0..n_values), 1), n x (n_values + 2) elements, in the
// caller's element format; the hash itself is launch_poseidon with arity n_values + 2
cudaError_t launch_smt_leaf_rows(const u32* keys, const u32* values, int n_values, size_t n, u32* rows, int mont,
                                 cudaStream_t stream);
// items covered by one resident wave of smt_path_kernel / varbase_window_kernel on the current device
size_t smt_path_wave_items(int sm_count);
size_t varbase_wave_items(int sm_count);
// arbo packed siblings (absolute offsets into a blob whose byte `base` is packed[0]) -> dense rows + bad[n]
// drop_is_old0 / drop_fnc1 (both or neither): arbo's post-insert rule, the last unpacked sibling is dropped where both are 0
cudaError_t launch_smt_unpack(const u8* packed, const u64* offsets, u64 base, u64 packed_bytes, size_t n, int n_levels,
                              u32* siblings, u8* bad, int mont, cudaStream_t stream, const u8* drop_is_old0 = nullptr,
                              const u8* drop_fnc1 = nullptr);
cudaError_t launch_smt_apply_bad(const u8* bad, size_t n, u8* flags, u8* status, u32* out_roots, cudaStream_t stream);

// ElGamal (elgamal.cuh)
// A fixed-base table is a 128-byte header {window bits, windows, entries per window} followed by the Niels entries; kernels
// take the pointer to the entries (fb_table_entries) and read the width from the header, so tables of different widths
// can be in use side by side.
constexpr int FB_TABLE_HEADER_WORDS = 32;
inline uint32_t* fb_table_entries(uint32_t* d_table) { return d_table ? d_table + FB_TABLE_HEADER_WORDS : nullptr; }
size_t fb_table_bytes(int wbits);
size_t fb_small_scratch_bytes();
cudaError_t upload_generator(u32* d_xy, cudaStream_t stream);
// te (last argument of the point-carrying launches below): the points on the wire are in iden3 twisted-Edwards coordinates
// (GCP_COORDS_TE): converted with one multiply on load / store (ecc/format/twistededwards.go:29-48)
cudaError_t launch_fb_table_build(const u32* d_base_xy, int base_mont, u32* d_small, u32* d_table, u32* d_flag,
                                  cudaStream_t stream, int te, int wbits);
cudaError_t launch_fixed_base_mul(const u32* tabG, const u32* scalars, size_t n, u32* out_xyz, u8* status, int mont,
                                  cudaStream_t stream);
cudaError_t launch_encrypt_shared(const u32* tabG, const u32* tabPK, const u32* pk_flag, const u32* ks, const u32* ms,
                                  size_t n, u32* out_xyz, u8* status, int mont, cudaStream_t stream);
// Encrypt with per-item keys, AssertDecrypt, DecryptionProof.Verify, EdDSA (varbase.cuh): several kernels per call on
// `stream`; scratch = varbase_scratch_bytes(kind, n) bytes (kind 0..3 in that order); *n_launches is incremented
size_t varbase_scratch_bytes(int kind, size_t n);
// out: n x 32 words (X, Y, Z, T) at scratch_result(...); the caller normalises with xyz_words = 32
size_t scalar_mul_scratch_bytes(size_t n, int n_bases);
cudaError_t launch_scalar_mul(const u32* points, const u32* scalars, const u32* points2, const u32* scalars2, size_t n,
                              u32* out_ext, u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream,
                              int te = 0);
cudaError_t launch_encrypt_per_key(const u32* tabG, const u32* pks, const u32* ks, const u32* ms, size_t n, u32* out_xyz,
                                   u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te = 0);
cudaError_t launch_normalize(const u32* xyz, size_t n_points, u32* out, u8* status, int pts_per_item, int mont,
                             cudaStream_t stream, int xyz_words = 24, int te = 0);
cudaError_t launch_ct_add(const u32* a, const u32* b, size_t n, u32* out_xyz, u8* status, int mont, cudaStream_t stream,
                          int te = 0);
cudaError_t launch_ct_neg(const u32* a, size_t n, u32* out, u8* status, cudaStream_t stream);
cudaError_t launch_ct_is_equal(const u32* a, const u32* b, size_t n, u8* flags, u8* status, cudaStream_t stream);
cudaError_t launch_ct_select(const u8* sel, const u32* i1, const u32* i2, size_t n, u32* out, u8* status, cudaStream_t stream);
cudaError_t launch_encrypt_tally(const u32* tabG, const u32* tabPK, const u32* ks, const u32* ms, const u8* mask,
                                 size_t n_ballots, int n_fields, int n_blocks, u32* partials, u32* bad_count, u32* out_xyz, u8* status, int mont,
                                 cudaStream_t stream, int m_words = 8);
cudaError_t launch_tally_status_merge(const u8* part_status, int n_chunks, int n_fields, int have_final, const u32* pk_flag,
                                      u32* ct, u8* status, cudaStream_t stream);
cudaError_t launch_keccak_address(const u8* in, size_t n, u8* out, cudaStream_t stream);
cudaError_t launch_assert_decrypt(const u32* tabG, const u32* cts, const u32* privs, const u32* msgs, size_t n, u8* flags,
                                  u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te = 0);
cudaError_t launch_decryption_proof(const u32* tabG, const PoseidonTable& tab13, const u32* pks, const u32* cts,
                                    const u32* msgs, const u32* a1s, const u32* a2s, const u32* zs, size_t n, u8* flags,
                                    u8* status, int mont, u32* scratch, int* n_launches, cudaStream_t stream, int te = 0);
cudaError_t launch_te_rte(const u32* in, size_t n_points, u32* out, u8* status, int to_rte, cudaStream_t stream);
cudaError_t launch_eddsa_verify(const u32* tabG, const PoseidonTable& tab6, const u32* pub_a, const u32* sig_r,
                                const u32* sig_s, const u32* msgs, size_t n, u8* flags, u8* status, int mont, u32* scratch,
                                int* n_launches, cudaStream_t stream);
cudaError_t upload_mimc7_constants(const u32* d_mont, cudaStream_t stream);
cudaError_t launch_mimc7(const u32* in, int len, size_t n, u32* out, u8* status, int mont, cudaStream_t stream);
cudaError_t launch_poseidon2_hash(const u32* keys, const u32* in, int len, size_t n, u32* out, u8* status, int mont,
                                  cudaStream_t stream);
cudaError_t launch_poseidon2_permutation(const u32* keys, const u32* in, size_t n, u32* out, u8* status, int mont,
                                         cudaStream_t stream);
int tally_max_blocks(size_t n_ballots, int n_fields, int sm_count);
cudaError_t launch_tally(const u32* ct, size_t n_ballots, int n_fields, int n_blocks, u32* partials, u32* bad_count,
                         u32* out_xyz, u8* status, int mont, cudaStream_t stream, int te = 0);

}  // namespace gcp

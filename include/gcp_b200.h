/* gcp_b200.h — C ABI of the B200 batch engine for the data-parallel core of
 * vocdoni/gnark-crypto-primitives.  This is the drop-in boundary: what a Go (cgo) caller, the
 * C++ mirror in gcp_b200.hpp and the Python ctypes mirror bind.  The reference has no FFI of its
 * own; every entry point below cites the Go gadget / helper whose VALUES it reproduces bit-exactly.
 *
 * Conventions
 *  - Field element (BN254 Fr): 32 bytes, little-endian.  GCP_FMT_CANONICAL: the integer itself, < r.
 *    GCP_FMT_MONTGOMERY: gnark-crypto fr.Element memory ([4]uint64 limbs, Montgomery, R = 2^256), so a
 *    Go caller can pass unsafe.Pointer(&elems[0]) of a []fr.Element unchanged.
 *  - Point: X then Y (64 bytes).  Ciphertext: C1.X, C1.Y, C2.X, C2.Y (128 bytes), the order of
 *    (*Ciphertext).Serialize, elgamal/ciphertext.go:98-105.
 *  - All arrays are flat and caller-owned; the engine never keeps a host pointer after returning
 *    (cgo pointer rules).  `*_dev` variants take device pointers and a cudaStream_t (as void*), enqueue
 *    work and return without synchronising; the plain variants take host pointers and block.
 *  - Return value: 0 on success, negative gcp_error on an engine fault (bad argument, CUDA error);
 *    gcp_last_error() gives the message.  A per-item `status` byte is non-zero where the reference
 *    would have failed an ASSERTION (solver error), while `flag` outputs carry the gadget's 0/1 result;
 *    an invalid proof is flag 0 / status 0, never an error.
 *  - Thread safety: calls on one gcp_ctx are serialised internally (a host-buffer call holds the context for its
 *    whole duration); use one ctx per GPU.  The `*_dev` variants return before their kernels have run and some of
 *    them keep intermediate results in per-context scratch (SMT leaf hashes and sort keys, projective points before
 *    normalisation, tally partials): work enqueued on one context through `*_dev` calls must therefore be ordered -
 *    the same stream, or events between streams - or spread over one context per stream.
 *  - There is no CPU fallback: without a CUDA device gcp_ctx_create fails.
 */
#ifndef GCP_B200_H
#define GCP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcp_ctx gcp_ctx;

enum gcp_error {
  GCP_OK = 0,
  GCP_ERR_BAD_ARG = -1,
  GCP_ERR_CUDA = -2,
  GCP_ERR_NO_DEVICE = -3,
  GCP_ERR_CONSTANTS = -4,
  GCP_ERR_ALLOC = -5
};

enum gcp_format { GCP_FMT_CANONICAL = 0, GCP_FMT_MONTGOMERY = 1 };
/* Or-ed into `fmt` of the gcp_elgamal_* / gcp_ballot_batch* / gcp_group_elgamal_* entry points (every call that reads or
 * writes curve points, EdDSA excepted: its points are TE by definition): the points on the wire are in iden3 / circom
 * twisted-Edwards coordinates, (x_TE, y) with x_RTE = x_TE * (-f) (FromTEtoRTE / FromRTEtoTE,
 * ecc/format/twistededwards.go:29-48).  They are converted inside the kernels (one multiply per point read or written),
 * so the call equals te_to_rte on every input point, the RTE call, rte_to_te on every output point - without the two extra
 * passes over PCIe.  On-curve assertions apply to the converted point.  Scalars and SMT elements are unaffected. */
enum gcp_coords { GCP_COORDS_RTE = 0, GCP_COORDS_TE = 2 };
/* Or-ed into the format argument of the fused tallies (gcp_elgamal_encrypt_tally*, gcp_ballot_batch*, their group forms):
 * the messages m are little-endian uint64 values, 8 bytes each, instead of 32-byte field elements (ballot fields are small
 * integers; the scalars k stay field elements).  40 instead of 64 bytes per encryption cross PCIe, which is what bounds
 * these calls once several GPUs are fed from one host.  The messages are integers in either element format. */
enum gcp_msg { GCP_MSG_FR = 0, GCP_MSG_U64 = 4 };

/* per-item status bytes */
enum gcp_status {
  GCP_STATUS_OK = 0,
  GCP_STATUS_NONCANONICAL = 1, /* an input element >= r */
  GCP_STATUS_KEY_RANGE = 2,    /* SMT key >= 2^n_levels: lowBits/ToBinary assertion, tree/smt/utils.go:11-13 */
  GCP_STATUS_NOT_BOOLEAN = 3,  /* enabled / fnc / isOld0 outside {0,1}: api.Select / api.And assertions */
  GCP_STATUS_OFF_CURVE = 4,    /* AssertIsOnCurve(pubKey), elgamal/encrypt.go:49 */
  GCP_STATUS_ZERO_DENOM = 5,   /* Edwards addition denominator 0 (reachable only with off-curve inputs) */
  GCP_STATUS_ASSERTION = 6,    /* an AssertIsEqual of the gadget fails (SMT processor, decryption checks) */
  GCP_STATUS_MALFORMED = 7     /* arbo.UnpackSiblings would reject the packed proof (tree/smt/wrapper_arbo.go:64-67) */
};

/* ---- context ------------------------------------------------------------------------------------ */
int gcp_device_count(void);
/* Loads the Poseidon tables (constants_path == NULL: the blob next to the library, data/poseidon_bn254.bin),
 * builds the fixed-base table of G on the device, creates streams and staging pools.  A context keeps about 2.2 GB
 * of device memory for the ElGamal tables (G and the cached election key: 13 windows x 2^19 entries x 96 B = 654 MB
 * each, plus the build scratch) besides its grow-only staging buffers; creation takes ~0.2 s. */
int gcp_ctx_create(int device, const char* constants_path, gcp_ctx** out);
void gcp_ctx_destroy(gcp_ctx* ctx);
const char* gcp_last_error(const gcp_ctx* ctx); /* ctx may be NULL: error of the last failed gcp_ctx_create */
int gcp_ctx_device(const gcp_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t gcp_ctx_launch_count(const gcp_ctx* ctx);

/* Host threads of the process-wide staging pool (csrc/hostcopy.h): min(16, hardware threads / 2), shared by every context
 * of the process; 0 before the first context exists.  GCP_B200_COPY_THREADS overrides it at first use. */
int gcp_copy_threads(void);
/* Measured bandwidth (GB/s) of that staging path on this host: `bytes` (>= 1 MB) of pageable memory copied into page-locked
 * memory by the pool.  A host-buffer call whose kernels outrun it is bound by this figure, not by the GPU. */
int gcp_copy_probe(size_t bytes, double* gb_per_s);
/* Page-locked host memory for the caller's flat arrays (cudaHostAlloc, portable across the devices of a group).
 * The host-buffer entry points accept any host pointer; from pinned memory their chunked copies run at full PCIe
 * rate and overlap the kernels (bench.py's e2e figure); from pageable memory (a Go heap slice) copies of 64 MB and more
 * are staged through the library's own page-locked ring by several host threads (about 4 % slower than page-locked
 * sources).  A Go caller wraps the pointer with unsafe.Slice.  Free with gcp_host_free. */
int gcp_host_alloc(size_t bytes, void** out);
void gcp_host_free(void* p);

/* Measures this GPU's integer-multiply pipe: sustained 32x32+64 multiply-adds per second (IMAD.WIDE.U32 carry
 * chains, the multiplier's own row primitive, no memory traffic).  bench.py uses it as the roofline denominator. */
int gcp_probe_imad_wide(gcp_ctx* ctx, double* wide_mul_per_s);

/* ---- Poseidon: hash/native/bn254/poseidon/poseidon.go ------------------------------------------- */
/* poseidon.Hash (poseidon.go:38-45, Sum :116-183): n independent hashes of `arity` inputs each.
 * in: n*arity elements, out: n elements.  arity outside 1..16 -> GCP_ERR_BAD_ARG ("bad inputs provided").
 * Host-buffer calls with at most 256 KB of inputs (BASELINE config 1: 1024 x 2) are latency-shaped: the kernel reads
 * and writes page-locked memory mapped into the device (one memcpy each way inside the library, no staged copies). */
int gcp_poseidon_hash(gcp_ctx* ctx, const void* in, int arity, size_t n, void* out, uint8_t* status, int fmt);
int gcp_poseidon_hash_dev(gcp_ctx* ctx, const void* d_in, int arity, size_t n, void* d_out, uint8_t* d_status,
                          int fmt, void* stream);
/* poseidon.MultiHash (poseidon.go:54-91): n independent multi-hashes of `len` inputs each, 1 <= len <= 4096. */
int gcp_poseidon_multihash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt);
int gcp_poseidon_multihash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status,
                               int fmt, void* stream);

/* MiMC7 (hash/native/bn254/mimc7/mimc.go:47-87): iden3-compatible, 91 rounds of x^7, Miyaguchi-Preneel with field
 * addition; n independent hashes of `len` inputs each, 1 <= len <= 62 (mimc.go:9). */
int gcp_mimc7_hash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt);
int gcp_mimc7_hash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status, int fmt,
                       void* stream);

/* Poseidon2, width 2 (hash/native/bn254/poseidon2): the Merkle-Damgard hasher HashPoseidon2.Hash (native.go:30-63) and
 * its gadget HashPoseidon2Gnark (gnark.go:18-54) over perm2 = poseidon2.NewPermutation(2, 6, 50) (native.go:27).
 * len = 2: internal node, the two limbs are ordered (min, max) first (native.go:42-44, MinMaxHint hints.go:10-19);
 * len = 3: leaf (key, value, flag); any other len -> GCP_ERR_BAD_ARG ("need 2 or 3 limbs", native.go:31-33).
 * CV_0 = 0, CV_{i+1} = Permutation([CV_i, m_i])[1] + m_i, out = CV_len.  In GCP_FMT_CANONICAL a 32-byte value >= r is
 * reduced mod r as SafeBigInt does (native.go:37-39; status stays 0); in GCP_FMT_MONTGOMERY it is status 1.
 * The permutation lives in gnark-crypto (un-vendored) and the reference holds no vector for it: the 62 round keys
 * (3 full rounds x 2, 50 partial rounds x 1, 3 full rounds x 2, in round order) are DATA.  A context starts with the
 * keys of data/poseidon2_bn254_t2.bin (the published Keccak-chain derivation restated by oracle/poseidon2.py, parity
 * unpinned); a Go host installs poseidon2.NewParameters(2, 6, 50).RoundKeys verbatim with gcp_poseidon2_set_round_keys
 * (n_keys must be 62, every key < r) and can compare gcp_poseidon2_permutation with perm2.Permutation at start-up. */
int gcp_poseidon2_set_round_keys(gcp_ctx* ctx, const void* keys, size_t n_keys, int fmt);
int gcp_poseidon2_hash(gcp_ctx* ctx, const void* in, int len, size_t n, void* out, uint8_t* status, int fmt);
int gcp_poseidon2_hash_dev(gcp_ctx* ctx, const void* d_in, int len, size_t n, void* d_out, uint8_t* d_status, int fmt,
                           void* stream);
/* perm2.Permutation on n states: in/out n x 2 elements. */
int gcp_poseidon2_permutation(gcp_ctx* ctx, const void* in, size_t n, void* out, uint8_t* status, int fmt);
int gcp_poseidon2_permutation_dev(gcp_ctx* ctx, const void* d_in, size_t n, void* d_out, uint8_t* d_status, int fmt,
                                  void* stream);

/* ---- SMT: tree/smt/verifier.go ------------------------------------------------------------------ */
/* The tree's hash function is a plug in the reference: `type Hasher func(frontend.API, ...frontend.Variable)`
 * (utils/hashers.go:10), handed to every gadget of tree/smt as hFn.  A context carries the plug for all of its
 * gcp_smt_* / gcp_ballot_batch* calls:
 *   GCP_HASHER_POSEIDON   utils.PoseidonHasher (:25-27): circomlib Poseidon, the hash of arbo's circom-compatible trees.  Default.
 *   GCP_HASHER_POSEIDON2  utils.Poseidon2Hasher (:35-37) = HashPoseidon2Gnark (hash/native/bn254/poseidon2/gnark.go:18-54):
 *                         the width-2 Merkle-Damgard hasher; a node hashes its children ordered (min, max), a leaf hashes
 *                         (key, value, 1); leaves with other than one value are "need 2 or 3 limbs" errors.  Uses the
 *                         context's round keys (gcp_poseidon2_set_round_keys; parity of the default key blob unpinned).
 * utils.MiMCHasher (:15-23) is gnark's std/hash/mimc, which lives outside the reference tree and is not implemented. */
enum gcp_hasher { GCP_HASHER_POSEIDON = 0, GCP_HASHER_POSEIDON2 = 1 };
int gcp_ctx_set_smt_hasher(gcp_ctx* ctx, int hasher);
int gcp_ctx_smt_hasher(const gcp_ctx* ctx);

/* smt.Verifier (verifier.go:102-121) -> VerifierWithLeafHashFlag (:171-242), n proofs of n_levels siblings.
 *   roots: n elements, or 1 element when shared_root != 0
 *   siblings: n * n_levels elements, root -> leaf, zero padded (Assignment.Siblings, wrapper.go:20-31)
 *   old_keys/old_values: n elements each (NULL,NULL => old leaf == new leaf, the InclusionVerifier form)
 *   is_old0, fnc, enabled: n bytes each or NULL (defaults 0, 0, 1)
 *   out_flags, out_status: n bytes each;  out_roots: n elements or NULL (the recomputed level[0])
 * n_levels must be in [2, 253]. */
int gcp_smt_verify(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                   const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* keys,
                   const void* values, const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags,
                   uint8_t* out_status, void* out_roots, int fmt);
int gcp_smt_verify_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_roots, int shared_root,
                       const void* d_siblings, const void* d_old_keys, const void* d_old_values,
                       const uint8_t* d_is_old0, const void* d_keys, const void* d_values, const uint8_t* d_fnc,
                       const uint8_t* d_enabled, uint8_t* d_out_flags, uint8_t* d_out_status, void* d_out_roots,
                       int fmt, void* stream);
/* smt.VerifierWithLeafHash / VerifierWithLeafHashFlag (verifier.go:129-183): the caller supplies the leaf hashes, so a
 * tree whose leaves carry several values (Hash1(key, values..., 1), tree/smt/hash.go:10-19) goes through the same path.
 * Arguments as gcp_smt_verify with hash1_old / hash1_new (n elements each) in the place of old_values / values;
 * old_keys and hash1_old are both given or both NULL (NULL,NULL: the old leaf is the new leaf).  The flag is the
 * gadget's; VerifierWithLeafHash itself asserts it (the caller's AssertIsEqual(valid, 1)). */
int gcp_smt_verify_with_leaf_hash(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                                  const void* siblings, const void* old_keys, const void* hash1_old, const uint8_t* is_old0,
                                  const void* keys, const void* hash1_new, const uint8_t* fnc, const uint8_t* enabled,
                                  uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt);
int gcp_smt_verify_with_leaf_hash_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_roots, int shared_root,
                                      const void* d_siblings, const void* d_old_keys, const void* d_hash1_old,
                                      const uint8_t* d_is_old0, const void* d_keys, const void* d_hash1_new,
                                      const uint8_t* d_fnc, const uint8_t* d_enabled, uint8_t* d_out_flags,
                                      uint8_t* d_out_status, void* d_out_roots, int fmt, void* stream);
/* smt.Hash1 (tree/smt/hash.go:10-19): out[i] = Poseidon(keys[i], values[i][0 .. n_values), 1), the leaf hash of a leaf
 * with n_values values (0 .. 14: the hasher takes at most 16 inputs, "bad inputs provided" beyond).  values: n x n_values
 * elements (may be NULL when n_values == 0).  n_values = 1 is the leaf of smt.Verifier / smt.Processor. */
int gcp_smt_leaf_hash(gcp_ctx* ctx, const void* keys, const void* values, int n_values, size_t n, void* out, uint8_t* status,
                      int fmt);
int gcp_smt_leaf_hash_dev(gcp_ctx* ctx, const void* d_keys, const void* d_values, int n_values, size_t n, void* d_out,
                          uint8_t* d_status, int fmt, void* stream);
/* The proof-streaming pass of the verifier on its own (LevInsFlag's inputs, lev_ins.go:43-77): per proof
 * lidx = 1 + index of the last non-zero sibling among [0, n-2] (0 if none) and info (bit 0: siblings[n-1] == 0,
 * bit 1: every sibling < r).  HBM-bound: reads n * n_levels * 32 bytes once with coalesced 128-bit loads. */
int gcp_smt_scan_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_siblings, uint16_t* d_lidx, uint8_t* d_info,
                     void* stream);
/* smt.InclusionVerifier (verifier.go:29-43). */
int gcp_smt_verify_inclusion(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                             const void* siblings, const void* keys, const void* values, uint8_t* out_flags,
                             uint8_t* out_status, void* out_roots, int fmt);
/* smt.ExclusionVerifier (verifier.go:66-81). */
int gcp_smt_verify_exclusion(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                             const void* siblings, const void* old_keys, const void* old_values,
                             const uint8_t* is_old0, const void* keys, uint8_t* out_flags, uint8_t* out_status,
                             void* out_roots, int fmt);

/* The same verifier fed with arbo's PACKED proofs, i.e. the byte strings GenProof returns, which the reference's
 * callers expand on the CPU with arbo.UnpackSiblings and pad with zeros to `levels` before they can fill
 * smt.Assignment (tree/smt/wrapper_arbo.go:48,63-76; testutil/utils.go:143,152-166).  Here the expansion runs on
 * the GPU, so a census-like proof crosses PCIe as ~0.8 KB instead of n_levels * 32 bytes.
 *   packed:  the n byte strings back to back;  offsets: n + 1 byte offsets into `packed` (offsets[i+1] - offsets[i]
 *            is the length of proof i)
 *   wire format of one proof (arbo PackSiblings):  u16 LE total length | u16 LE bitmap length L | L bitmap bytes
 *            (bit i = byte i/8, bit i%8: sibling i is non-zero) | 32 bytes LE per set bit
 *   siblings are canonical little-endian integers on the wire in either `fmt`; every other element follows `fmt`.
 * A string arbo.UnpackSiblings would reject gets status GCP_STATUS_MALFORMED, flag 0; siblings past n_levels are
 * dropped exactly as wrapper_arbo.go:69-76 drops them.  Everything else as gcp_smt_verify. */
int gcp_smt_verify_packed(gcp_ctx* ctx, int n_levels, size_t n, const void* roots, int shared_root,
                          const uint8_t* packed, const uint64_t* offsets, const void* old_keys, const void* old_values,
                          const uint8_t* is_old0, const void* keys, const void* values, const uint8_t* fnc,
                          const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status, void* out_roots, int fmt);
/* arbo.UnpackSiblings + zero padding on device buffers: d_siblings receives n * n_levels elements in `fmt`,
 * d_bad n bytes (0 or GCP_STATUS_MALFORMED).  Compose with gcp_smt_verify_dev / gcp_smt_process_dev. */
int gcp_smt_unpack_siblings_dev(gcp_ctx* ctx, int n_levels, size_t n, const uint8_t* d_packed, size_t packed_bytes,
                                const uint64_t* d_offsets, void* d_siblings, uint8_t* d_bad, int fmt, void* stream);

/* smt.Processor (tree/smt/processor.go:10-72): state transition of n independent trees/proofs.
 * fnc = (fnc0, fnc1): (1,0) insert, (0,1) update, (1,1) delete, (0,0) nop.  Output: new_roots (n elements).
 * The gadget ASSERTS (old root matches the proof, LevIns, final state, update keeps the key, isOld0 boolean), so a
 * failing item gets status 6 (or 2/3/1) and new_root = 0. */
int gcp_smt_process(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const void* siblings,
                    const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* new_keys,
                    const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1, void* new_roots, uint8_t* status,
                    int fmt);
/* The same with arbo packed proofs (wrapper_arbo.go:166-179 unpacks them on the CPU): see gcp_smt_verify_packed for
 * the wire format; a string arbo.UnpackSiblings would reject gets status GCP_STATUS_MALFORMED and new_root = 0.
 * Two forms, which differ in WHEN the proof was generated:
 *   gcp_smt_process_packed  the unpacked siblings are used as they are: proofs taken BEFORE the change (GenProof on
 *                           the old tree), or any caller that has already applied arbo's post-insert rule;
 *   gcp_smt_process_arbo    the reference's own flow, addOrUpdate (wrapper_arbo.go:152-172): GenProof runs AFTER the
 *                           add, so for an insert beside an existing leaf the last unpacked sibling is the displaced
 *                           old leaf and is dropped: where is_old0[i] == 0 and fnc1[i] == 0 the last sibling of the
 *                           string is left out (a string with no sibling to drop gets GCP_STATUS_MALFORMED: the
 *                           reference's slice expression panics there).  Feeding post-insert strings to
 *                           gcp_smt_process_packed puts the insertion level one too deep and every such insert
 *                           fails its old-root assertion (status 6). */
int gcp_smt_process_arbo(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const uint8_t* packed,
                         const uint64_t* offsets, const void* old_keys, const void* old_values, const uint8_t* is_old0,
                         const void* new_keys, const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1,
                         void* new_roots, uint8_t* status, int fmt);
int gcp_smt_process_packed(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const uint8_t* packed,
                           const uint64_t* offsets, const void* old_keys, const void* old_values, const uint8_t* is_old0,
                           const void* new_keys, const void* new_values, const uint8_t* fnc0, const uint8_t* fnc1,
                           void* new_roots, uint8_t* status, int fmt);
int gcp_smt_process_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_old_roots, const void* d_siblings,
                        const void* d_old_keys, const void* d_old_values, const uint8_t* d_is_old0,
                        const void* d_new_keys, const void* d_new_values, const uint8_t* d_fnc0, const uint8_t* d_fnc1,
                        void* d_new_roots, uint8_t* d_status, int fmt, void* stream);
/* smt.ProcessorWithLeafHash (processor.go:16-72): hash1_old / hash1_new in the place of old_values / new_values. */
int gcp_smt_process_with_leaf_hash(gcp_ctx* ctx, int n_levels, size_t n, const void* old_roots, const void* siblings,
                                   const void* old_keys, const void* hash1_old, const uint8_t* is_old0,
                                   const void* new_keys, const void* hash1_new, const uint8_t* fnc0, const uint8_t* fnc1,
                                   void* new_roots, uint8_t* status, int fmt);
int gcp_smt_process_with_leaf_hash_dev(gcp_ctx* ctx, int n_levels, size_t n, const void* d_old_roots, const void* d_siblings,
                                       const void* d_old_keys, const void* d_hash1_old, const uint8_t* d_is_old0,
                                       const void* d_new_keys, const void* d_hash1_new, const uint8_t* d_fnc0,
                                       const uint8_t* d_fnc1, void* d_new_roots, uint8_t* d_status, int fmt, void* stream);

/* ---- ElGamal over the a = -1 BN254 twisted Edwards curve: elgamal/ ------------------------------------ */
/* Multiplications by G and by the shared public key run out of precomputed window tables in device memory (the reference's
 * table, elgamal/mul.go:26-72, has 4-bit windows).  A table starts with 20-bit windows (654 MB, 13 additions per
 * multiplication) and is rebuilt with 24-bit windows (8.9 GB, 11 additions, ~45 ms) once its base has served 2^27
 * multiplications; results do not depend on the width.  gcp_ctx_set_fixed_base_window fixes the width of both tables
 * (8..26 bits; 0 restores the automatic choice): G's table is rebuilt at once, the key's at its next use.
 * gcp_ctx_fixed_base_window returns the current width of G's (which = 0) or the cached key's (which = 1) table.  For a
 * group, call it on every device's context (gcp_group_ctx). */
int gcp_ctx_set_fixed_base_window(gcp_ctx* ctx, int window_bits);
int gcp_ctx_fixed_base_window(const gcp_ctx* ctx, int which);
/* FixedBaseScalarMulBN254 (elgamal/mul.go:76-166): out[i] = [scalars[i]] G, scalars are Fr elements used as
 * integers in [0, r).  out_points: n x (X, Y). */
int gcp_elgamal_fixed_base_mul(gcp_ctx* ctx, const void* scalars, size_t n, void* out_points, uint8_t* status, int fmt);
int gcp_elgamal_fixed_base_mul_dev(gcp_ctx* ctx, const void* d_scalars, size_t n, void* d_out_points,
                                   uint8_t* d_status, int fmt, void* stream);
/* curve.ScalarMul of gnark's twistededwards gadget over a batch (the variable-base multiplication the reference calls at
 * elgamal/encrypt.go:55, elgamal/ciphertext.go:58,147-160, ecc/bn254/eddsa/verifier.go:71-80): out[i] = [scalars[i]] points[i]
 * for scalars used as integers in [0, r).  With points2 / scalars2 (both or neither): out[i] = [scalars[i]] points[i] +
 * [scalars2[i]] points2[i] in ONE pass that shares the doublings (Straus), the form DecryptionProof.Verify's
 * z*C1 - e*D takes.  Points must be on the curve (status 4; the reference's callers assert it first: encrypt.go:49,
 * ciphertext.go:53-54,131-135); points with a cofactor component are multiplied by the integer scalar.
 * points, points2, out_points: n x (X, Y). */
int gcp_elgamal_scalar_mul(gcp_ctx* ctx, const void* points, const void* scalars, const void* points2,
                           const void* scalars2, size_t n, void* out_points, uint8_t* status, int fmt);
int gcp_elgamal_scalar_mul_dev(gcp_ctx* ctx, const void* d_points, const void* d_scalars, const void* d_points2,
                               const void* d_scalars2, size_t n, void* d_out_points, uint8_t* d_status, int fmt,
                               void* stream);
/* (*Ciphertext).Encrypt (elgamal/encrypt.go:42-64): C1 = [k]G, C2 = [m]G + [k]pubKey.  m = 0 gives EncryptedZero
 * (encrypt.go:72-94).  pub_key: one point (pk_per_item = 0, the election key) or n points (pk_per_item = 1).  The
 * shared-key form builds a 654 MB window table for the key on first use (~0.1 s, cached in the context until another
 * key is used): it is meant for many ballots under one key; for a handful of encryptions per key use the per-item
 * form (windowed variable-base multiplication, no table).  status 4 where AssertIsOnCurve(pubKey) would fail.
 * out_ct: n x (C1.X, C1.Y, C2.X, C2.Y). */
int gcp_elgamal_encrypt(gcp_ctx* ctx, const void* pub_key, int pk_per_item, const void* k, const void* m, size_t n,
                        void* out_ct, uint8_t* status, int fmt);
int gcp_elgamal_encrypt_dev(gcp_ctx* ctx, const void* d_pub_key, int pk_per_item, const void* d_k, const void* d_m,
                            size_t n, void* d_out_ct, uint8_t* d_status, int fmt, void* stream);
/* (*Ciphertext).Add (elgamal/ciphertext.go:24-32), element-wise over n ciphertext pairs; inputs are not checked to
 * be on the curve (as in the reference); status 5 where an addition denominator is zero. */
int gcp_elgamal_add(gcp_ctx* ctx, const void* a, const void* b, size_t n, void* out, uint8_t* status, int fmt);
int gcp_elgamal_add_dev(gcp_ctx* ctx, const void* d_a, const void* d_b, size_t n, void* d_out, uint8_t* d_status,
                        int fmt, void* stream);
/* (*Ciphertext).Neg (elgamal/ciphertext.go:37-46). */
int gcp_elgamal_neg(gcp_ctx* ctx, const void* a, size_t n, void* out, uint8_t* status, int fmt);
/* (*Ciphertext).IsEqual (elgamal/ciphertext.go:79-87): out_flags[i] = 1 iff the four coordinates of a[i] and b[i]
 * agree; (*Ciphertext).Select (:90-96): out[i] = sel[i] ? i1[i] : i2[i], status 3 where sel[i] is not 0/1 (api.Select
 * asserts a boolean).  Both work on either element format (they compare / copy canonical representations). */
int gcp_elgamal_is_equal(gcp_ctx* ctx, const void* a, const void* b, size_t n, uint8_t* out_flags, uint8_t* status);
int gcp_elgamal_select(gcp_ctx* ctx, const uint8_t* sel, const void* i1, const void* i2, size_t n, void* out,
                       uint8_t* status);
/* Tally: per field f, the fold of Ciphertext.Add over ct[0..n_ballots)[f] starting from NewCiphertext
 * (ciphertext.go:16-32).  ct: n_ballots x n_fields ciphertexts; out: n_fields ciphertexts; status: n_fields bytes.
 * Inputs must be curve points (outputs of Encrypt); the reduction order is unspecified, which is exact for group
 * elements.  Multi-GPU: tally each shard, all-gather the n_fields partial ciphertexts, tally the gathered array. */
int gcp_elgamal_tally(gcp_ctx* ctx, const void* ct, size_t n_ballots, int n_fields, void* out, uint8_t* status, int fmt);
int gcp_elgamal_tally_dev(gcp_ctx* ctx, const void* d_ct, size_t n_ballots, int n_fields, void* d_out,
                          uint8_t* d_status, int fmt, void* stream);

/* Fused Encrypt + tally: out[f] = sum over ballots of Encrypt(pub_key, k[b][f], m[b][f]) without materialising
 * the ciphertexts (the config-3 shape: 2^24 ballots x 8 fields would be 17 GB of ciphertexts).  k, m: n_ballots x
 * n_fields scalars; out: n_fields ciphertexts; status: n_fields bytes (4 everywhere if the key is off the curve,
 * 1 for a field that saw a non-canonical scalar; such items are left out of the sum). */
int gcp_elgamal_encrypt_tally(gcp_ctx* ctx, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                              int n_fields, void* out, uint8_t* status, int fmt);
int gcp_elgamal_encrypt_tally_dev(gcp_ctx* ctx, const void* d_pub_key, const void* d_k, const void* d_m,
                                  size_t n_ballots, int n_fields, void* d_out, uint8_t* d_status, int fmt,
                                  void* stream);

/* (*Ciphertext).AssertDecrypt (elgamal/ciphertext.go:50-67): flag[i] = 1 iff C1, C2 are on the curve (else status 4)
 * and C2 - [priv]C1 == [m]G. */
int gcp_elgamal_assert_decrypt(gcp_ctx* ctx, const void* ct, const void* priv_keys, const void* msgs, size_t n,
                               uint8_t* out_flags, uint8_t* status, int fmt);
/* DecryptionProof.Verify (elgamal/ciphertext.go:124-168) with hFn = poseidon.MultiHash: flag[i] = 1 iff
 * z*G == A1 + e*P and z*C1 == A2 + e*D, D = C2 - [msg]G, e = MultiHash(P, P, C1, D, A1, A2); status 4 if any of
 * P, C1, C2, A1, A2 is off the curve.  pub_keys, a1, a2: n points; ct: n ciphertexts; msgs, z: n elements. */
int gcp_elgamal_verify_decryption_proof(gcp_ctx* ctx, const void* pub_keys, const void* ct, const void* msgs,
                                        const void* a1, const void* a2, const void* z, size_t n, uint8_t* out_flags,
                                        uint8_t* status, int fmt);
/* EdDSA-Poseidon Verifier.IsValid (ecc/bn254/eddsa/verifier.go:55-88): public key A and signature point R in
 * circom/iden3 TE coordinates, S already reduced mod the subgroup order (types.go:37-49).  flag[i] = the gadget's
 * result; status 4 where PointToRTE's AssertIsOnCurve would fail. */
int gcp_eddsa_verify(gcp_ctx* ctx, const void* pub_keys_te, const void* sig_r_te, const void* sig_s, const void* msgs,
                     size_t n, uint8_t* out_flags, uint8_t* status, int fmt);
/* format.FromTEtoRTE / FromRTEtoTE (ecc/format/twistededwards.go:29-48): x' = x * (-f) resp. x / (-f), y unchanged,
 * between circom/iden3 BabyJubJub (a = 168700) and gnark's a = -1 form.  Works in either element format. */
int gcp_te_to_rte(gcp_ctx* ctx, const void* points, size_t n_points, void* out, uint8_t* status);
int gcp_rte_to_te(gcp_ctx* ctx, const void* points, size_t n_points, void* out, uint8_t* status);

/* ---- End-to-end ballot batch (BASELINE config 5) ------------------------------------------------------- */
/* Per voter v: flag[v] = smt.InclusionVerifier(census proof v) (tree/smt/verifier.go:29-43); the tally is the fold of
 * Ciphertext.Add (elgamal/ciphertext.go:24-32) over Encrypt(pub_key, k[v][f], m[v][f]) (elgamal/encrypt.go:42-64) of
 * the voters with flag 1.  Device buffers only (2^26 voters x 5.2 KB of proofs do not fit host or one GPU: callers
 * stream chunks and tally the per-chunk results with gcp_elgamal_tally_dev). */
int gcp_ballot_batch_dev(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* d_roots, int shared_root,
                         const void* d_siblings, const void* d_keys, const void* d_values, const void* d_pub_key,
                         const void* d_k, const void* d_m, int n_fields, uint8_t* d_flags, uint8_t* d_status,
                         void* d_tally, uint8_t* d_tally_status, int fmt, void* stream);

/* Host-buffer form: voters are streamed to the device in chunks (copy of chunk k+1 overlaps the kernels of chunk
 * k).  Census proofs are dense rows (`siblings`) or, with siblings == NULL, arbo packed strings (`packed`, `offsets`:
 * see gcp_smt_verify_packed).  out_flags / out_status: n_voters bytes each; out_tally: n_fields ciphertexts,
 * out_tally_status: n_fields bytes. */
int gcp_ballot_batch(gcp_ctx* ctx, int n_levels, size_t n_voters, const void* roots, int shared_root, const void* siblings,
                     const uint8_t* packed, const uint64_t* offsets, const void* keys, const void* values,
                     const void* pub_key, const void* k, const void* m, int n_fields, uint8_t* out_flags,
                     uint8_t* out_status, void* out_tally, uint8_t* out_tally_status, int fmt);

/* ---- Ethereum address: ecc/secp256k1/ecdsa/address.go:14-40 ------------------------------------------- */
/* DeriveAddress: out_addr[i] = Keccak256_legacy(pub_xy_be[i])[12:32], pub_xy_be[i] = X (32 bytes big-endian) ||
 * Y (32 bytes big-endian).  out_addr: n x 20 bytes (the bytes U8ToVar packs big-endian into one variable). */
int gcp_keccak_address(gcp_ctx* ctx, const void* pub_xy_be, size_t n, void* out_addr);
int gcp_keccak_address_dev(gcp_ctx* ctx, const void* d_pub_xy_be, size_t n, void* d_out_addr, void* stream);

/* ---- several GPUs of one box behind one handle (single-process callers: the Go host) -------------------------- */
/* The reference has no multi-device code; its work units are independent (every gadget call touches only its own
 * inputs), so a group shards each batch by contiguous index range [n*i/g, n*(i+1)/g) over its g devices, one host
 * thread and one context per device, with no data-path collective.  The one exchange step is the ElGamal tally:
 * every device folds its slice to n_fields partial ciphertexts, the partials (n_fields * 128 bytes per device) are
 * all-gathered as bytes with ncclAllGather over NVLink / NVSwitch, and every device folds the gathered array; the
 * result is bit-identical for any device count.  NCCL is bound at run time (libnccl.so.2); a group of more than one
 * device cannot be created without it.  Calls on one group are serialised; per-item results land at the caller's
 * global index. */
typedef struct gcp_group gcp_group;
int gcp_group_create(const int* devices, int n_devices, const char* constants_path, gcp_group** out);
void gcp_group_destroy(gcp_group* g);
const char* gcp_group_last_error(const gcp_group* g); /* g may be NULL: error of the last failed gcp_group_create */
int gcp_group_size(const gcp_group* g);
gcp_ctx* gcp_group_ctx(gcp_group* g, int i); /* device i's context, for every entry point that has no group form */
int gcp_group_uses_nccl(const gcp_group* g); /* 1 when the partial tallies travel through ncclAllGather (g > 1) */
/* sharded forms of gcp_poseidon_hash, gcp_smt_verify, gcp_smt_verify_packed, gcp_elgamal_encrypt (same arguments) */
int gcp_group_poseidon_hash(gcp_group* g, const void* in, int arity, size_t n, void* out, uint8_t* status, int fmt);
int gcp_group_smt_verify(gcp_group* g, int n_levels, size_t n, const void* roots, int shared_root, const void* siblings,
                         const void* old_keys, const void* old_values, const uint8_t* is_old0, const void* keys,
                         const void* values, const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags,
                         uint8_t* out_status, void* out_roots, int fmt);
int gcp_group_smt_verify_packed(gcp_group* g, int n_levels, size_t n, const void* roots, int shared_root,
                                const uint8_t* packed, const uint64_t* offsets, const void* old_keys,
                                const void* old_values, const uint8_t* is_old0, const void* keys, const void* values,
                                const uint8_t* fnc, const uint8_t* enabled, uint8_t* out_flags, uint8_t* out_status,
                                void* out_roots, int fmt);
int gcp_group_elgamal_encrypt(gcp_group* g, const void* pub_key, int pk_per_item, const void* k, const void* m, size_t n,
                              void* out_ct, uint8_t* status, int fmt);
/* sharded tally / fused encrypt + tally with the all-gather of the partial ciphertexts (arguments as the ctx forms) */
int gcp_group_elgamal_tally(gcp_group* g, const void* ct, size_t n_ballots, int n_fields, void* out, uint8_t* status,
                            int fmt);
int gcp_group_elgamal_encrypt_tally(gcp_group* g, const void* pub_key, const void* k, const void* m, size_t n_ballots,
                                    int n_fields, void* out, uint8_t* status, int fmt);
/* gcp_ballot_batch over the voters sharded across the group (BASELINE config 5 at 1/2/4/8 GPUs); the per-device
 * tallies are exchanged like gcp_group_elgamal_tally's. */
int gcp_group_ballot_batch(gcp_group* g, int n_levels, size_t n_voters, const void* roots, int shared_root,
                           const void* siblings, const uint8_t* packed, const uint64_t* offsets, const void* keys,
                           const void* values, const void* pub_key, const void* k, const void* m, int n_fields,
                           uint8_t* out_flags, uint8_t* out_status, void* out_tally, uint8_t* out_tally_status, int fmt);

#ifdef __cplusplus
}
#endif
#endif /* GCP_B200_H */

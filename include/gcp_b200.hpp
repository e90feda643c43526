// gcp_b200.hpp — header-only C++ mirror of the reference's Go-facing names over the C ABI (gcp_b200.h).
// Same names, argument meaning and error behaviour as vocdoni/gnark-crypto-primitives, but batched:
//   poseidon::Hash / MultiHash        hash/native/bn254/poseidon/poseidon.go:38,54
//   smt::InclusionVerifier / ExclusionVerifier / Verifier / Processor   tree/smt/verifier.go:29,66,102, processor.go:10
//   elgamal::Encrypt / Add / Neg / Tally / FixedBaseScalarMulBN254       elgamal/encrypt.go:42, ciphertext.go:24,37, mul.go:76
// Errors of the reference ("bad inputs provided", ...) surface as gcp::Error; per-item assertion failures as status
// bytes; flags as the gadget's 0/1 result.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include "gcp_b200.h"

namespace gcp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

using Element = std::array<uint8_t, 32>;  // little-endian; canonical integer or gnark-crypto Montgomery memory

class Engine {
 public:
  explicit Engine(int device = 0, const char* constants_path = nullptr) {
    int rc = gcp_ctx_create(device, constants_path, &ctx_);
    if (rc != GCP_OK) throw Error(rc, gcp_last_error(nullptr));
  }
  ~Engine() { gcp_ctx_destroy(ctx_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  gcp_ctx* raw() const { return ctx_; }
  void check(int rc) const {
    if (rc != GCP_OK) throw Error(rc, gcp_last_error(ctx_));
  }

 private:
  gcp_ctx* ctx_ = nullptr;
};

struct Batch {
  std::vector<uint8_t> values;  // n x (elements) x 32 bytes
  std::vector<uint8_t> status;  // n
  std::vector<uint8_t> flags;   // n (verifiers only)
};

namespace poseidon {
// n rows of `arity` inputs -> n digests.  arity outside 1..16 throws "bad inputs provided" (poseidon.go:41-43).
inline Batch Hash(const Engine& e, const uint8_t* inputs, int arity, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_poseidon_hash(e.raw(), inputs, arity, n, b.values.data(), b.status.data(), fmt));
  return b;
}
inline Batch MultiHash(const Engine& e, const uint8_t* inputs, int len, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_poseidon_multihash(e.raw(), inputs, len, n, b.values.data(), b.status.data(), fmt));
  return b;
}
// Batched mirror of the reference's stateful hasher (hash.Hash[T], hash/hash.go:9-18; poseidon.go:94-197): every
// Write carries one column (n elements, one per row of the batch); a Write that would exceed 16 inputs in total is
// dropped whole and silently (poseidon.go:103-108); Sum runs all rows on the GPU.
class Hasher {
 public:
  Hasher(const Engine& e, size_t n_rows, int fmt = GCP_FMT_CANONICAL) : e_(e), n_(n_rows), fmt_(fmt) {}
  void Write(std::initializer_list<const uint8_t*> columns) {
    if (cols_.size() + columns.size() > 16) return;
    for (const uint8_t* c : columns) cols_.emplace_back(c, c + n_ * 32);
  }
  void Reset() { cols_.clear(); }
  bool WriteSucceeded() const { return !cols_.empty(); }
  Batch Sum() const {
    std::vector<uint8_t> rows(n_ * cols_.size() * 32);
    for (size_t i = 0; i < n_; i++)
      for (size_t j = 0; j < cols_.size(); j++)
        std::copy(cols_[j].begin() + i * 32, cols_[j].begin() + (i + 1) * 32, rows.begin() + (i * cols_.size() + j) * 32);
    return Hash(e_, rows.data(), (int)cols_.size(), n_, fmt_);  // no input: "bad inputs provided"
  }
  // flags[i] = 1 iff Sum()[i] == expected[i]
  Batch SumIsEqual(const uint8_t* expected) const {
    Batch b = Sum();
    b.flags.resize(n_);
    for (size_t i = 0; i < n_; i++)
      b.flags[i] = b.status[i] == 0 && std::equal(b.values.begin() + i * 32, b.values.begin() + (i + 1) * 32, expected + i * 32);
    return b;
  }
  void AssertSumIsEqual(const uint8_t* expected) const {
    Batch b = SumIsEqual(expected);
    for (size_t i = 0; i < n_; i++)
      if (!b.flags[i]) throw Error(GCP_ERR_BAD_ARG, "AssertSumIsEqual failed for row " + std::to_string(i));
  }

 private:
  const Engine& e_;
  size_t n_;
  int fmt_;
  std::vector<std::vector<uint8_t>> cols_;
};
}  // namespace poseidon

namespace poseidon2 {
// HashPoseidon2.Hash / HashPoseidon2Gnark (hash/native/bn254/poseidon2/native.go:30-63, gnark.go:18-54): n rows of 2
// limbs (internal node, ordered min/max) or 3 limbs (leaf); any other count throws "need 2 or 3 limbs".
inline Batch Hash(const Engine& e, const uint8_t* limbs, int len, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_poseidon2_hash(e.raw(), limbs, len, n, b.values.data(), b.status.data(), fmt));
  return b;
}
// perm2.Permutation (native.go:27,55) on n states of two elements.
inline Batch Permutation(const Engine& e, const uint8_t* states, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 64);
  b.status.resize(n);
  e.check(gcp_poseidon2_permutation(e.raw(), states, n, b.values.data(), b.status.data(), fmt));
  return b;
}
// Installs gnark-crypto's own round keys (62 elements in round order); see the header.
inline void SetRoundKeys(const Engine& e, const uint8_t* keys, size_t n_keys, int fmt = GCP_FMT_CANONICAL) {
  e.check(gcp_poseidon2_set_round_keys(e.raw(), keys, n_keys, fmt));
}
}  // namespace poseidon2

namespace smt {
inline Batch InclusionVerifier(const Engine& e, int n_levels, size_t n, const uint8_t* roots, bool shared_root,
                               const uint8_t* siblings, const uint8_t* keys, const uint8_t* values,
                               int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.flags.resize(n);
  b.status.resize(n);
  e.check(gcp_smt_verify_inclusion(e.raw(), n_levels, n, roots, shared_root, siblings, keys, values, b.flags.data(),
                                   b.status.data(), nullptr, fmt));
  return b;
}
inline Batch ExclusionVerifier(const Engine& e, int n_levels, size_t n, const uint8_t* roots, bool shared_root,
                               const uint8_t* siblings, const uint8_t* old_keys, const uint8_t* old_values,
                               const uint8_t* is_old0, const uint8_t* keys, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.flags.resize(n);
  b.status.resize(n);
  e.check(gcp_smt_verify_exclusion(e.raw(), n_levels, n, roots, shared_root, siblings, old_keys, old_values, is_old0, keys,
                                   b.flags.data(), b.status.data(), nullptr, fmt));
  return b;
}
inline Batch Verifier(const Engine& e, int n_levels, size_t n, const uint8_t* enabled, const uint8_t* roots, bool shared_root,
                      const uint8_t* siblings, const uint8_t* old_keys, const uint8_t* old_values, const uint8_t* is_old0,
                      const uint8_t* keys, const uint8_t* values, const uint8_t* fnc, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.flags.resize(n);
  b.status.resize(n);
  e.check(gcp_smt_verify(e.raw(), n_levels, n, roots, shared_root, siblings, old_keys, old_values, is_old0, keys, values, fnc,
                         enabled, b.flags.data(), b.status.data(), nullptr, fmt));
  return b;
}
// Verifier over arbo's packed proofs (GenProof output, expanded on the GPU instead of arbo.UnpackSiblings + padding,
// tree/smt/wrapper_arbo.go:63-76); old_keys == nullptr gives the InclusionVerifier form.
inline Batch VerifierPacked(const Engine& e, int n_levels, size_t n, const uint8_t* roots, bool shared_root,
                            const uint8_t* packed, const uint64_t* offsets, const uint8_t* old_keys,
                            const uint8_t* old_values, const uint8_t* is_old0, const uint8_t* keys, const uint8_t* values,
                            const uint8_t* fnc, const uint8_t* enabled = nullptr, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.flags.resize(n);
  b.status.resize(n);
  e.check(gcp_smt_verify_packed(e.raw(), n_levels, n, roots, shared_root, packed, offsets, old_keys, old_values, is_old0,
                                keys, values, fnc, enabled, b.flags.data(), b.status.data(), nullptr, fmt));
  return b;
}
// returns the new roots in `values`
inline Batch Processor(const Engine& e, int n_levels, size_t n, const uint8_t* old_roots, const uint8_t* siblings,
                       const uint8_t* old_keys, const uint8_t* old_values, const uint8_t* is_old0, const uint8_t* new_keys,
                       const uint8_t* new_values, const uint8_t* fnc0, const uint8_t* fnc1, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_smt_process(e.raw(), n_levels, n, old_roots, siblings, old_keys, old_values, is_old0, new_keys, new_values, fnc0,
                          fnc1, b.values.data(), b.status.data(), fmt));
  return b;
}
// window width of the precomputed tables behind FixedBaseScalarMulBN254 / Encrypt (elgamal/mul.go:26-72): 8..26, 0 = automatic
inline void SetFixedBaseWindow(const Engine& e, int window_bits) { e.check(gcp_ctx_set_fixed_base_window(e.raw(), window_bits)); }
// the hFn utils.Hasher argument of the tree/smt gadgets (utils/hashers.go:10-37): carried by the engine
inline void SetHasher(const Engine& e, int hasher /* GCP_HASHER_POSEIDON | GCP_HASHER_POSEIDON2 */) {
  e.check(gcp_ctx_set_smt_hasher(e.raw(), hasher));
}
// smt.ProcessorWithLeafHash (processor.go:16): hash1_old / hash1_new in the place of the values
inline Batch ProcessorWithLeafHash(const Engine& e, int n_levels, size_t n, const uint8_t* old_roots, const uint8_t* siblings,
                                   const uint8_t* old_keys, const uint8_t* hash1_old, const uint8_t* is_old0,
                                   const uint8_t* new_keys, const uint8_t* hash1_new, const uint8_t* fnc0, const uint8_t* fnc1,
                                   int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_smt_process_with_leaf_hash(e.raw(), n_levels, n, old_roots, siblings, old_keys, hash1_old, is_old0, new_keys,
                                         hash1_new, fnc0, fnc1, b.values.data(), b.status.data(), fmt));
  return b;
}
// smt.Processor fed as WrapperArbo.addOrUpdate feeds it (wrapper_arbo.go:152-172): packed proofs generated AFTER the change
inline Batch ProcessorArbo(const Engine& e, int n_levels, size_t n, const uint8_t* old_roots, const uint8_t* packed,
                           const uint64_t* offsets, const uint8_t* old_keys, const uint8_t* old_values, const uint8_t* is_old0,
                           const uint8_t* new_keys, const uint8_t* new_values, const uint8_t* fnc0, const uint8_t* fnc1,
                           int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_smt_process_arbo(e.raw(), n_levels, n, old_roots, packed, offsets, old_keys, old_values, is_old0, new_keys,
                               new_values, fnc0, fnc1, b.values.data(), b.status.data(), fmt));
  return b;
}
// smt.Hash1 (hash.go:10-19): H(key, values..., 1), n_values values per leaf; digests in `values`
inline Batch Hash1(const Engine& e, const uint8_t* keys, const uint8_t* values, int n_values, size_t n,
                   int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 32);
  b.status.resize(n);
  e.check(gcp_smt_leaf_hash(e.raw(), keys, values, n_values, n, b.values.data(), b.status.data(), fmt));
  return b;
}
// smt.VerifierWithLeafHashFlag (verifier.go:171-242); VerifierWithLeafHash (:129-147) asserts the flag
inline Batch VerifierWithLeafHashFlag(const Engine& e, int n_levels, size_t n, const uint8_t* enabled, const uint8_t* roots,
                                      bool shared_root, const uint8_t* siblings, const uint8_t* old_keys,
                                      const uint8_t* hash1_old, const uint8_t* is_old0, const uint8_t* keys,
                                      const uint8_t* hash1_new, const uint8_t* fnc, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.flags.resize(n);
  b.status.resize(n);
  e.check(gcp_smt_verify_with_leaf_hash(e.raw(), n_levels, n, roots, shared_root, siblings, old_keys, hash1_old, is_old0, keys,
                                        hash1_new, fnc, enabled, b.flags.data(), b.status.data(), nullptr, fmt));
  return b;
}
}  // namespace smt

namespace elgamal {
// ciphertexts are 4 elements in Serialize() order: C1.X, C1.Y, C2.X, C2.Y (ciphertext.go:98-105)
inline Batch Encrypt(const Engine& e, const uint8_t* pub_key, bool pk_per_item, const uint8_t* k, const uint8_t* m, size_t n,
                     int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 128);
  b.status.resize(n);
  e.check(gcp_elgamal_encrypt(e.raw(), pub_key, pk_per_item, k, m, n, b.values.data(), b.status.data(), fmt));
  return b;
}
inline Batch FixedBaseScalarMulBN254(const Engine& e, const uint8_t* scalars, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 64);
  b.status.resize(n);
  e.check(gcp_elgamal_fixed_base_mul(e.raw(), scalars, n, b.values.data(), b.status.data(), fmt));
  return b;
}
inline Batch Add(const Engine& e, const uint8_t* x, const uint8_t* y, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 128);
  b.status.resize(n);
  e.check(gcp_elgamal_add(e.raw(), x, y, n, b.values.data(), b.status.data(), fmt));
  return b;
}
inline Batch Neg(const Engine& e, const uint8_t* x, size_t n, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize(n * 128);
  b.status.resize(n);
  e.check(gcp_elgamal_neg(e.raw(), x, n, b.values.data(), b.status.data(), fmt));
  return b;
}
inline Batch Tally(const Engine& e, const uint8_t* ct, size_t n_ballots, int n_fields, int fmt = GCP_FMT_CANONICAL) {
  Batch b;
  b.values.resize((size_t)n_fields * 128);
  b.status.resize(n_fields);
  e.check(gcp_elgamal_tally(e.raw(), ct, n_ballots, n_fields, b.values.data(), b.status.data(), fmt));
  return b;
}
}  // namespace elgamal

// Several GPUs of one box behind one handle (gcp_group_*): index-range shards, NCCL all-gather of the partial tallies.
class Group {
 public:
  explicit Group(const std::vector<int>& devices, const char* constants_path = nullptr) {
    int rc = gcp_group_create(devices.data(), (int)devices.size(), constants_path, &grp_);
    if (rc != GCP_OK) throw Error(rc, gcp_group_last_error(nullptr));
  }
  ~Group() { gcp_group_destroy(grp_); }
  Group(const Group&) = delete;
  Group& operator=(const Group&) = delete;
  gcp_group* raw() const { return grp_; }
  int size() const { return gcp_group_size(grp_); }
  void check(int rc) const {
    if (rc != GCP_OK) throw Error(rc, gcp_group_last_error(grp_));
  }
  Batch InclusionVerifier(int n_levels, size_t n, const uint8_t* roots, bool shared_root, const uint8_t* siblings,
                          const uint8_t* keys, const uint8_t* values, int fmt = GCP_FMT_CANONICAL) const {
    Batch b;
    b.flags.resize(n);
    b.status.resize(n);
    check(gcp_group_smt_verify(grp_, n_levels, n, roots, shared_root, siblings, nullptr, nullptr, nullptr, keys, values,
                               nullptr, nullptr, b.flags.data(), b.status.data(), nullptr, fmt));
    return b;
  }
  Batch Tally(const uint8_t* ct, size_t n_ballots, int n_fields, int fmt = GCP_FMT_CANONICAL) const {
    Batch b;
    b.values.resize((size_t)n_fields * 128);
    b.status.resize(n_fields);
    check(gcp_group_elgamal_tally(grp_, ct, n_ballots, n_fields, b.values.data(), b.status.data(), fmt));
    return b;
  }

 private:
  gcp_group* grp_ = nullptr;
};

}  // namespace gcp

// C++ caller of the C ABI through the header-only mirror (include/gcp_b200.hpp): Poseidon([1,2]) known answer,
// a three-level inclusion proof built with the engine's own hashes, and the reference's error text for a bad arity.
#include <cstdio>
#include <cstring>

#include "gcp_b200.hpp"

static void put(uint8_t* dst, uint64_t v) {
  memset(dst, 0, 32);
  memcpy(dst, &v, 8);
}

int main() {
  gcp::Engine eng(0);
  // Poseidon([1, 2]) = 7853200120776062878684798364095072458815029376092732009249414926327459813530 (circomlib vector)
  uint8_t in[64];
  put(in, 1);
  put(in + 32, 2);
  gcp::Batch h = gcp::poseidon::Hash(eng, in, 2, 1);
  static const uint8_t want[32] = {0x9a, 0x18, 0x17, 0x44, 0x7a, 0x60, 0x19, 0x9e, 0x51, 0x45, 0x32, 0x74, 0xf2, 0x17, 0x36, 0x2a,
                                   0xcf, 0xe9, 0x62, 0x96, 0x6b, 0x4c, 0xf6, 0x3d, 0x41, 0x90, 0xd6, 0xe7, 0xf5, 0xc0, 0x5c, 0x11};
  if (h.status[0] != 0 || memcmp(h.values.data(), want, 32) != 0) {
    fprintf(stderr, "poseidon mismatch\n");
    return 1;
  }
  // inclusion proof over 3 levels with siblings (11, 22, 0), key 7, value 9: root = H(22, H(11, H(7, 9, 1)))... built via the engine
  uint8_t leaf_in[96], node_in[64], sib[96], key[32], val[32];
  put(leaf_in, 7); put(leaf_in + 32, 9); put(leaf_in + 64, 1);
  gcp::Batch leaf = gcp::poseidon::Hash(eng, leaf_in, 3, 1);
  put(sib, 11); put(sib + 32, 22); put(sib + 64, 0);
  // key 7 = bits 1,1,1: at level 1 the path is on the right: H(sib1, leaf); at level 0: H(sib0, that)
  memcpy(node_in, sib + 32, 32); memcpy(node_in + 32, leaf.values.data(), 32);
  gcp::Batch n1 = gcp::poseidon::Hash(eng, node_in, 2, 1);
  memcpy(node_in, sib, 32); memcpy(node_in + 32, n1.values.data(), 32);
  gcp::Batch root = gcp::poseidon::Hash(eng, node_in, 2, 1);
  put(key, 7); put(val, 9);
  gcp::Batch v = gcp::smt::InclusionVerifier(eng, 3, 1, root.values.data(), false, sib, key, val);
  if (v.flags[0] != 1 || v.status[0] != 0) {
    fprintf(stderr, "inclusion proof rejected\n");
    return 2;
  }
  put(key, 5);  // tree/smt/utils_test.go:27-39: key 5 must not verify where key 7 does
  v = gcp::smt::InclusionVerifier(eng, 3, 1, root.values.data(), false, sib, key, val);
  if (v.flags[0] != 0 || v.status[0] != 0) return 3;
  // the same proof in arbo's packed form: total length 4 + 1 + 2*32, bitmap length 1, bitmap 0b011, siblings 11, 22
  uint8_t packed[4 + 1 + 64] = {69, 0, 1, 0, 3};
  memcpy(packed + 5, sib, 64);
  const uint64_t offsets[2] = {0, sizeof(packed)};
  put(key, 7);
  v = gcp::smt::VerifierPacked(eng, 3, 1, root.values.data(), false, packed, offsets, nullptr, nullptr, nullptr, key, val,
                               nullptr);
  if (v.flags[0] != 1 || v.status[0] != 0) {
    fprintf(stderr, "packed inclusion proof rejected\n");
    return 6;
  }
  packed[0] = 70;  // length field mismatch: arbo.UnpackSiblings errors
  v = gcp::smt::VerifierPacked(eng, 3, 1, root.values.data(), false, packed, offsets, nullptr, nullptr, nullptr, key, val,
                               nullptr);
  if (v.flags[0] != 0 || v.status[0] != GCP_STATUS_MALFORMED) return 7;
  try {
    gcp::poseidon::Hash(eng, in, 17, 1);
    return 4;
  } catch (const gcp::Error& e) {
    if (std::string(e.what()).find("bad inputs provided") == std::string::npos) return 5;
  }
  {  // the same proof through a one-device group
    gcp::Group grp({0});
    put(key, 7);
    gcp::Batch gv = grp.InclusionVerifier(3, 1, root.values.data(), false, sib, key, val);
    if (grp.size() != 1 || gv.flags[0] != 1 || gv.status[0] != 0) return 8;
  }
  {  // stateful hasher mirror: Write / Sum / SumIsEqual (hash/hash.go:9-18)
    uint8_t c1[32], c2[32];
    put(c1, 1); put(c2, 2);
    gcp::poseidon::Hasher hs(eng, 1);
    if (hs.WriteSucceeded()) return 9;
    hs.Write({c1, c2});
    gcp::Batch d = hs.Sum();
    uint8_t two[64];
    memcpy(two, c1, 32); memcpy(two + 32, c2, 32);
    gcp::Batch ref = gcp::poseidon::Hash(eng, two, 2, 1);
    if (!hs.WriteSucceeded() || d.values != ref.values || hs.SumIsEqual(ref.values.data()).flags[0] != 1) return 10;
    const uint8_t* seventeen[17];
    for (int i = 0; i < 17; i++) seventeen[i] = c1;
    hs.Write({c1, c1, c1, c1, c1, c1, c1, c1, c1, c1, c1, c1, c1, c1, c1});  // 2 + 15 > 16: dropped whole
    if (hs.Sum().values != ref.values) return 11;
    (void)seventeen;
  }
  {  // width-2 Poseidon2 hasher: a node is ordered (min, max), a leaf is not; wrong arity throws the reference's text
    uint8_t ab[64], ba[64], leaf_ab[96], leaf_ba[96], one[32];
    put(ab, 5); put(ab + 32, 9); put(ba, 9); put(ba + 32, 5); put(one, 1);
    memcpy(leaf_ab, ab, 64); memcpy(leaf_ab + 64, one, 32);
    memcpy(leaf_ba, ba, 64); memcpy(leaf_ba + 64, one, 32);
    gcp::Batch n1 = gcp::poseidon2::Hash(eng, ab, 2, 1), n2 = gcp::poseidon2::Hash(eng, ba, 2, 1);
    gcp::Batch l1 = gcp::poseidon2::Hash(eng, leaf_ab, 3, 1), l2 = gcp::poseidon2::Hash(eng, leaf_ba, 3, 1);
    if (n1.status[0] != 0 || n1.values != n2.values || l1.values == l2.values) return 12;
    gcp::Batch pm = gcp::poseidon2::Permutation(eng, ab, 1);
    if (pm.values.size() != 64 || pm.status[0] != 0) return 13;
    try {
      gcp::poseidon2::Hash(eng, ab, 4, 1);
      return 14;
    } catch (const gcp::Error& e) {
      if (std::string(e.what()).find("need 2 or 3 limbs") == std::string::npos) return 15;
    }
  }
  printf("cpp mirror ok\n");
  return 0;
}

"""CPU oracle for the B200 batch engine — TEST INFRASTRUCTURE ONLY.

A restatement, in plain Python integers (and in plain C under oracle/c/), of the
value semantics of the vocdoni/gnark-crypto-primitives gadgets on the hot path
(SURVEY.md section 8).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import or execute anything in this
directory; the product (gnark_crypto_primitives_b200) never does.

Parity pinning: the Go reference cannot run in this environment (no Go
toolchain, un-vendored modules), so the oracle is pinned to
  * the reference's own constant tables (oracle/gen_constants.py parses
    hash/native/bn254/poseidon/constants.go, sha256-checked),
  * the reference's only fully static known-answer test
    (elgamal/ciphertext_test.go:286-345, Chaum-Pedersen decryption proof), which
    exercises curve parameters, Add, Neg, ScalarMul, FixedBase and MultiHash t=13,
  * the scaling-factor constant of ecc/format/twistededwards.go:17,
  * the public circomlib Poseidon vectors and public Keccak/Ethereum vectors,
  * the reference's fixed-input self-consistency tests (encrypt_test.go:96-217,
    tree/smt/utils_test.go:27-39).
See tests/test_oracle_golden.py.
"""

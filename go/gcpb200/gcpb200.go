// Package gcpb200 is the cgo binding of the B200 batch engine (include/gcp_b200.h) for Go callers of
// vocdoni/gnark-crypto-primitives: witness generation and batch verification code can hand whole slices of
// fr.Element to the GPU and get back exactly the values the gadgets would constrain.
//
// NOTE: written against the header; it has never been compiled — there is no Go toolchain in the build
// environment (see INTEGRATION.md).  It is deliberately thin and mechanical: flat slices in, flat slices out,
// no Go pointer is retained by the C side (cgo pointer rules), every call blocks until results are written.
package gcpb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gnark_crypto_primitives_b200 -lgcp_b200 -Wl,-rpath,${SRCDIR}/../../gnark_crypto_primitives_b200
#include "gcp_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math/big"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/vocdoni/gnark-crypto-primitives/tree/smt"
)

// Status bytes (per item): non-zero where the reference gadget would fail an assertion.
const (
	StatusOK           = 0
	StatusNonCanonical = 1
	StatusKeyRange     = 2 // tree/smt/utils.go:11-13
	StatusNotBoolean   = 3
	StatusOffCurve     = 4 // elgamal/encrypt.go:49
	StatusZeroDenom    = 5
	StatusAssertion    = 6 // tree/smt/processor.go assertions
	StatusMalformed    = 7 // arbo.UnpackSiblings error (tree/smt/wrapper_arbo.go:64-67)
)

// Engine owns one GPU context. Goroutines may share it (calls are serialised inside the library); use one Engine
// per GPU for parallelism.
type Engine struct{ ctx *C.gcp_ctx }

// New creates a context on the given CUDA device. There is no CPU fallback: without a GPU this fails.
func New(device int) (*Engine, error) {
	var ctx *C.gcp_ctx
	// The creation error text is process-wide in the library (mutex-guarded, readable from any thread), so the two cgo
	// calls below may land on different OS threads; another goroutine failing a create in between can still replace it.
	if rc := C.gcp_ctx_create(C.int(device), nil, &ctx); rc != 0 {
		return nil, fmt.Errorf("gcp_ctx_create: %s", C.GoString(C.gcp_last_error(nil)))
	}
	return &Engine{ctx}, nil
}

func (e *Engine) Close() { C.gcp_ctx_destroy(e.ctx) }

func (e *Engine) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.gcp_last_error(e.ctx)))
}

func elemPtr(s []fr.Element) unsafe.Pointer {
	if len(s) == 0 {
		return nil
	}
	return unsafe.Pointer(&s[0]) // fr.Element is [4]uint64 Montgomery limbs: GCP_FMT_MONTGOMERY memory
}

func bytePtr(s []byte) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0]))
}

// BatchHash mirrors poseidon.Hash (hash/native/bn254/poseidon/poseidon.go:38) over len(in)/arity rows.
// arity outside 1..16 returns the reference's error "bad inputs provided".
func (e *Engine) BatchHash(in []fr.Element, arity int) (out []fr.Element, status []byte, err error) {
	if arity <= 0 || len(in)%arity != 0 {
		return nil, nil, errors.New("bad inputs provided")
	}
	n := len(in) / arity
	out, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_poseidon_hash(e.ctx, elemPtr(in), C.int(arity), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// BatchMultiHash mirrors poseidon.MultiHash (poseidon.go:54) over rows of `length` inputs (1..4096).
func (e *Engine) BatchMultiHash(in []fr.Element, length int) (out []fr.Element, status []byte, err error) {
	if length <= 0 || len(in)%length != 0 {
		return nil, nil, errors.New("bad inputs provided")
	}
	n := len(in) / length
	out, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_poseidon_multihash(e.ctx, elemPtr(in), C.int(length), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// InstallPoseidon2Keys hands gnark-crypto's own round keys of poseidon2.NewPermutation(2, 6, 50)
// (hash/native/bn254/poseidon2/native.go:27) to the engine and checks one permutation against gnark-crypto, so that the
// width-2 Poseidon2 hasher is pinned to the library the reference links at start-up rather than to a restated key
// derivation.  params is poseidon2.NewParameters(2, 6, 50) and perm is poseidon2.NewPermutation(2, 6, 50), both from
// github.com/consensys/gnark-crypto/ecc/bn254/fr/poseidon2.
func (e *Engine) InstallPoseidon2Keys(roundKeys [][]fr.Element, permute func([]fr.Element) error) error {
	flat := make([]fr.Element, 0, 62)
	for _, row := range roundKeys {
		flat = append(flat, row...)
	}
	if err := e.err(C.gcp_poseidon2_set_round_keys(e.ctx, elemPtr(flat), C.size_t(len(flat)), C.GCP_FMT_MONTGOMERY)); err != nil {
		return err
	}
	var in, want, got [2]fr.Element
	in[0].SetUint64(1)
	in[1].SetUint64(2)
	want = in
	if err := permute(want[:]); err != nil {
		return err
	}
	status := make([]byte, 1)
	if err := e.err(C.gcp_poseidon2_permutation(e.ctx, unsafe.Pointer(&in[0]), 1, unsafe.Pointer(&got[0]), bytePtr(status),
		C.GCP_FMT_MONTGOMERY)); err != nil {
		return err
	}
	if status[0] != 0 || !got[0].Equal(&want[0]) || !got[1].Equal(&want[1]) {
		return errors.New("gcpb200: Poseidon2 permutation differs from gnark-crypto's")
	}
	return nil
}

// BatchPoseidon2 mirrors HashPoseidon2.Hash / utils.Poseidon2Hasher (hash/native/bn254/poseidon2/native.go:30,
// utils/hashers.go:35) over len(in)/limbs rows: limbs = 2 is an internal node (ordered min, max), limbs = 3 a leaf
// (key, value, flag).
func (e *Engine) BatchPoseidon2(in []fr.Element, limbs int) (out []fr.Element, status []byte, err error) {
	if (limbs != 2 && limbs != 3) || len(in)%limbs != 0 {
		return nil, nil, fmt.Errorf("poseidon2: need 2 or 3 limbs, got %d", limbs)
	}
	n := len(in) / limbs
	out, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_poseidon2_hash(e.ctx, elemPtr(in), C.int(limbs), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// Proofs is the flattened form of []smt.Assignment (tree/smt/wrapper.go:20-31): Siblings holds n*Levels elements,
// root -> leaf, zero padded; Roots holds n elements or a single shared root.
type Proofs struct {
	Levels    int
	Roots     []fr.Element
	Siblings  []fr.Element
	OldKeys   []fr.Element // nil for plain inclusion proofs
	OldValues []fr.Element
	IsOld0    []byte
	Keys      []fr.Element
	Values    []fr.Element
	Fnc       []byte // nil: inclusion (0)
	Enabled   []byte // nil: 1
}

// BatchVerify mirrors smt.Verifier (tree/smt/verifier.go:102); with OldKeys == nil it is smt.InclusionVerifier (:29).
// flags[i] is the gadget's 0/1 result; status[i] != 0 means the gadget would have failed an assertion.
func (e *Engine) BatchVerify(p *Proofs) (flags, status []byte, err error) {
	n := len(p.Keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(p.Roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_smt_verify(e.ctx, C.int(p.Levels), C.size_t(n), elemPtr(p.Roots), C.int(shared), elemPtr(p.Siblings),
		elemPtr(p.OldKeys), elemPtr(p.OldValues), bytePtr(p.IsOld0), elemPtr(p.Keys), elemPtr(p.Values), bytePtr(p.Fnc),
		bytePtr(p.Enabled), bytePtr(flags), bytePtr(status), nil, C.GCP_FMT_MONTGOMERY))
	return
}

// BatchVerifyPacked is BatchVerify for proofs still in arbo's packed form: packed[i] is the siblingsPacked that
// tree.GenProof returned for Keys[i] (the reference's callers run arbo.UnpackSiblings and pad to `levels` on the CPU,
// tree/smt/wrapper_arbo.go:63-76); p.Siblings is ignored.  status 7 = arbo.UnpackSiblings would have failed.
func (e *Engine) BatchVerifyPacked(p *Proofs, packed [][]byte) (flags, status []byte, err error) {
	n := len(p.Keys)
	flags, status = make([]byte, n), make([]byte, n)
	offsets := make([]uint64, n+1)
	total := 0
	for i, b := range packed {
		total += len(b)
		offsets[i+1] = uint64(total)
	}
	blob := make([]byte, 0, total+1)
	for _, b := range packed {
		blob = append(blob, b...)
	}
	blob = append(blob, 0) // never a nil pointer for an all-empty batch
	shared := 0
	if len(p.Roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_smt_verify_packed(e.ctx, C.int(p.Levels), C.size_t(n), elemPtr(p.Roots), C.int(shared), bytePtr(blob),
		(*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(p.OldKeys), elemPtr(p.OldValues), bytePtr(p.IsOld0),
		elemPtr(p.Keys), elemPtr(p.Values), bytePtr(p.Fnc), bytePtr(p.Enabled), bytePtr(flags), bytePtr(status), nil,
		C.GCP_FMT_MONTGOMERY))
	return
}

// BatchProcess mirrors smt.Processor (tree/smt/processor.go:10): fnc (1,0) insert, (0,1) update, (1,1) delete, (0,0) nop.
func (e *Engine) BatchProcess(levels int, oldRoots, siblings, oldKeys, oldValues []fr.Element, isOld0 []byte,
	newKeys, newValues []fr.Element, fnc0, fnc1 []byte) (newRoots []fr.Element, status []byte, err error) {
	n := len(newKeys)
	newRoots, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_smt_process(e.ctx, C.int(levels), C.size_t(n), elemPtr(oldRoots), elemPtr(siblings), elemPtr(oldKeys),
		elemPtr(oldValues), bytePtr(isOld0), elemPtr(newKeys), elemPtr(newValues), bytePtr(fnc0), bytePtr(fnc1),
		elemPtr(newRoots), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// SetFixedBaseWindow fixes the window width (8..26 bits; 0 = automatic: 20 bits, 24 once a base has served 2^27
// multiplications) of the precomputed tables behind FixedBaseScalarMulBN254 and Encrypt (elgamal/mul.go:26-72 uses 4-bit
// windows).  Results do not depend on it; 24 bits cost 8.9 GB of device memory per base.
func (e *Engine) SetFixedBaseWindow(bits int) error {
	return e.err(C.gcp_ctx_set_fixed_base_window(e.ctx, C.int(bits)))
}

// SetSMTHasher selects the utils.Hasher plug (utils/hashers.go:10-37) for every SMT call of this engine: the gadgets of
// tree/smt take hFn per call, an Engine carries it.  poseidon2 = false: utils.PoseidonHasher (default); true:
// utils.Poseidon2Hasher (call InstallPoseidon2Keys first, so that the round keys are gnark-crypto's own).
func (e *Engine) SetSMTHasher(poseidon2 bool) error {
	h := C.int(C.GCP_HASHER_POSEIDON)
	if poseidon2 {
		h = C.int(C.GCP_HASHER_POSEIDON2)
	}
	return e.err(C.gcp_ctx_set_smt_hasher(e.ctx, h))
}

// BatchProcessWithLeafHash mirrors smt.ProcessorWithLeafHash (tree/smt/processor.go:16): the caller supplies
// hash1Old / hash1New (e.g. from BatchHash1 for leaves with several values).
func (e *Engine) BatchProcessWithLeafHash(levels int, oldRoots, siblings, oldKeys, hash1Old []fr.Element, isOld0 []byte,
	newKeys, hash1New []fr.Element, fnc0, fnc1 []byte) (newRoots []fr.Element, status []byte, err error) {
	n := len(newKeys)
	newRoots, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_smt_process_with_leaf_hash(e.ctx, C.int(levels), C.size_t(n), elemPtr(oldRoots), elemPtr(siblings),
		elemPtr(oldKeys), elemPtr(hash1Old), bytePtr(isOld0), elemPtr(newKeys), elemPtr(hash1New), bytePtr(fnc0), bytePtr(fnc1),
		elemPtr(newRoots), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// BatchProcessArbo is smt.Processor fed the way WrapperArbo.addOrUpdate feeds it (tree/smt/wrapper_arbo.go:152-172):
// packed[i] is the siblingsPacked of a GenProof taken AFTER the add/update; where isOld0[i] == 0 and fnc1[i] == 0 the last
// unpacked sibling (the displaced old leaf) is dropped, as :170-172 do, before the row is padded to `levels`.
func (e *Engine) BatchProcessArbo(levels int, oldRoots []fr.Element, packed [][]byte, oldKeys, oldValues []fr.Element,
	isOld0 []byte, newKeys, newValues []fr.Element, fnc0, fnc1 []byte) (newRoots []fr.Element, status []byte, err error) {
	n := len(newKeys)
	newRoots, status = make([]fr.Element, n), make([]byte, n)
	offsets := make([]uint64, n+1)
	total := 0
	for i, b := range packed {
		total += len(b)
		offsets[i+1] = uint64(total)
	}
	blob := make([]byte, 0, total+1)
	for _, b := range packed {
		blob = append(blob, b...)
	}
	blob = append(blob, 0)
	err = e.err(C.gcp_smt_process_arbo(e.ctx, C.int(levels), C.size_t(n), elemPtr(oldRoots), bytePtr(blob),
		(*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(oldKeys), elemPtr(oldValues), bytePtr(isOld0), elemPtr(newKeys),
		elemPtr(newValues), bytePtr(fnc0), bytePtr(fnc1), elemPtr(newRoots), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// BatchHash1 mirrors smt.Hash1 (tree/smt/hash.go:10-19): out[i] = Poseidon(keys[i], values[i*nValues : (i+1)*nValues]..., 1).
func (e *Engine) BatchHash1(keys, values []fr.Element, nValues int) (out []fr.Element, status []byte, err error) {
	n := len(keys)
	out, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_smt_leaf_hash(e.ctx, elemPtr(keys), elemPtr(values), C.int(nValues), C.size_t(n), elemPtr(out),
		bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// BatchVerifyWithLeafHash mirrors smt.VerifierWithLeafHashFlag (tree/smt/verifier.go:171): p.OldValues / p.Values carry
// hash1Old / hash1New.  smt.VerifierWithLeafHash (:129) is this plus the caller's check that every flag is 1.
func (e *Engine) BatchVerifyWithLeafHash(p *Proofs) (flags, status []byte, err error) {
	n := len(p.Keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(p.Roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_smt_verify_with_leaf_hash(e.ctx, C.int(p.Levels), C.size_t(n), elemPtr(p.Roots), C.int(shared),
		elemPtr(p.Siblings), elemPtr(p.OldKeys), elemPtr(p.OldValues), bytePtr(p.IsOld0), elemPtr(p.Keys), elemPtr(p.Values),
		bytePtr(p.Fnc), bytePtr(p.Enabled), bytePtr(flags), bytePtr(status), nil, C.GCP_FMT_MONTGOMERY))
	return
}

// BatchEncrypt mirrors (*Ciphertext).Encrypt (elgamal/encrypt.go:42) with one shared public key (X, Y).
// Ciphertexts come back as 4 elements each in Serialize() order: C1.X, C1.Y, C2.X, C2.Y (ciphertext.go:98-105).
func (e *Engine) BatchEncrypt(pubKey [2]fr.Element, k, m []fr.Element) (ct []fr.Element, status []byte, err error) {
	n := len(k)
	ct, status = make([]fr.Element, 4*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_encrypt(e.ctx, unsafe.Pointer(&pubKey[0]), 0, elemPtr(k), elemPtr(m), C.size_t(n), elemPtr(ct),
		bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// CiphertextAdd mirrors (*Ciphertext).Add (elgamal/ciphertext.go:24), element-wise over n pairs.
func (e *Engine) CiphertextAdd(a, b []fr.Element) (out []fr.Element, status []byte, err error) {
	n := len(a) / 4
	out, status = make([]fr.Element, 4*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_add(e.ctx, elemPtr(a), elemPtr(b), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// Tally folds Ciphertext.Add over ballots per field: ct is nBallots x nFields ciphertexts.
func (e *Engine) Tally(ct []fr.Element, nFields int) (out []fr.Element, status []byte, err error) {
	nBallots := len(ct) / (4 * nFields)
	out, status = make([]fr.Element, 4*nFields), make([]byte, nFields)
	err = e.err(C.gcp_elgamal_tally(e.ctx, elemPtr(ct), C.size_t(nBallots), C.int(nFields), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// EncryptTally encrypts nBallots x nFields values under one key and returns only the nFields aggregated ciphertexts.
func (e *Engine) EncryptTally(pubKey [2]fr.Element, k, m []fr.Element, nFields int) (out []fr.Element, status []byte, err error) {
	nBallots := len(k) / nFields
	out, status = make([]fr.Element, 4*nFields), make([]byte, nFields)
	err = e.err(C.gcp_elgamal_encrypt_tally(e.ctx, unsafe.Pointer(&pubKey[0]), elemPtr(k), elemPtr(m), C.size_t(nBallots),
		C.int(nFields), elemPtr(out), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// EncryptTallyU64 is EncryptTally with the messages as plain uint64 values (ballot fields are small integers): 40 instead
// of 64 bytes per encryption cross PCIe, which is what bounds the call once several GPUs are fed from one host.
func (e *Engine) EncryptTallyU64(pubKey [2]fr.Element, k []fr.Element, m []uint64, nFields int) (out []fr.Element, status []byte, err error) {
	nBallots := len(k) / nFields
	if len(m) != len(k) {
		return nil, nil, errors.New("EncryptTallyU64: one message per scalar")
	}
	out, status = make([]fr.Element, 4*nFields), make([]byte, nFields)
	var mp unsafe.Pointer
	if len(m) > 0 {
		mp = unsafe.Pointer(&m[0])
	}
	err = e.err(C.gcp_elgamal_encrypt_tally(e.ctx, unsafe.Pointer(&pubKey[0]), elemPtr(k), mp, C.size_t(nBallots),
		C.int(nFields), elemPtr(out), bytePtr(status), C.GCP_FMT_MONTGOMERY|C.GCP_MSG_U64))
	return
}

// DeriveAddresses mirrors ecdsa.DeriveAddress (ecc/secp256k1/ecdsa/address.go:14): pub is n x 64 bytes X_be||Y_be.
func (e *Engine) DeriveAddresses(pub []byte) (addr []byte, err error) {
	n := len(pub) / 64
	addr = make([]byte, 20*n)
	err = e.err(C.gcp_keccak_address(e.ctx, unsafe.Pointer(bytePtr(pub)), C.size_t(n), unsafe.Pointer(bytePtr(addr))))
	return
}

// packBlob lays byte strings back to back: the (blob, offsets) form of the packed-proof entry points.
func packBlob(packed [][]byte) (blob []byte, offsets []uint64) {
	offsets = make([]uint64, len(packed)+1)
	total := 0
	for i, b := range packed {
		total += len(b)
		offsets[i+1] = uint64(total)
	}
	blob = make([]byte, 0, total+1)
	for _, b := range packed {
		blob = append(blob, b...)
	}
	blob = append(blob, 0) // never a nil pointer for an all-empty batch
	return
}

// CoordsTE is or-ed into the format of the point-carrying calls below (fmtOf): points on the wire are in iden3 / circom
// twisted-Edwards coordinates and are converted inside the kernels (format.FromTEtoRTE / FromRTEtoTE,
// ecc/format/twistededwards.go:29-48).
type Coords int

const (
	CoordsRTE Coords = 0
	CoordsTE  Coords = Coords(C.GCP_COORDS_TE)
)

func fmtOf(c Coords) C.int { return C.int(C.GCP_FMT_MONTGOMERY) | C.int(c) }

// BatchInclusionVerify mirrors smt.InclusionVerifier (tree/smt/verifier.go:29).
func (e *Engine) BatchInclusionVerify(levels int, roots, siblings, keys, values []fr.Element) (flags, status []byte, err error) {
	n := len(keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_smt_verify_inclusion(e.ctx, C.int(levels), C.size_t(n), elemPtr(roots), C.int(shared), elemPtr(siblings),
		elemPtr(keys), elemPtr(values), bytePtr(flags), bytePtr(status), nil, C.GCP_FMT_MONTGOMERY))
	return
}

// BatchExclusionVerify mirrors smt.ExclusionVerifier (tree/smt/verifier.go:66).
func (e *Engine) BatchExclusionVerify(levels int, roots, siblings, oldKeys, oldValues []fr.Element, isOld0 []byte,
	keys []fr.Element) (flags, status []byte, err error) {
	n := len(keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_smt_verify_exclusion(e.ctx, C.int(levels), C.size_t(n), elemPtr(roots), C.int(shared), elemPtr(siblings),
		elemPtr(oldKeys), elemPtr(oldValues), bytePtr(isOld0), elemPtr(keys), bytePtr(flags), bytePtr(status), nil,
		C.GCP_FMT_MONTGOMERY))
	return
}

// BatchProcessPacked is BatchProcess over arbo packed proofs taken BEFORE the change (see BatchProcessArbo for the
// reference's own post-insert flow).
func (e *Engine) BatchProcessPacked(levels int, oldRoots []fr.Element, packed [][]byte, oldKeys, oldValues []fr.Element,
	isOld0 []byte, newKeys, newValues []fr.Element, fnc0, fnc1 []byte) (newRoots []fr.Element, status []byte, err error) {
	n := len(newKeys)
	newRoots, status = make([]fr.Element, n), make([]byte, n)
	blob, offsets := packBlob(packed)
	err = e.err(C.gcp_smt_process_packed(e.ctx, C.int(levels), C.size_t(n), elemPtr(oldRoots), bytePtr(blob),
		(*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(oldKeys), elemPtr(oldValues), bytePtr(isOld0), elemPtr(newKeys),
		elemPtr(newValues), bytePtr(fnc0), bytePtr(fnc1), elemPtr(newRoots), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// FixedBaseScalarMul mirrors FixedBaseScalarMulBN254 (elgamal/mul.go:76): points come back as (X, Y) pairs.
func (e *Engine) FixedBaseScalarMul(scalars []fr.Element, c Coords) (points []fr.Element, status []byte, err error) {
	n := len(scalars)
	points, status = make([]fr.Element, 2*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_fixed_base_mul(e.ctx, elemPtr(scalars), C.size_t(n), elemPtr(points), bytePtr(status), fmtOf(c)))
	return
}

// ScalarMul mirrors curve.ScalarMul (gnark twistededwards; elgamal/encrypt.go:55): out[i] = [scalars[i]] points[i], or with a
// second base [s]P + [s2]P2 in one pass (points2 / scalars2 both nil or both given).
func (e *Engine) ScalarMul(points, scalars, points2, scalars2 []fr.Element, c Coords) (out []fr.Element, status []byte, err error) {
	n := len(scalars)
	out, status = make([]fr.Element, 2*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_scalar_mul(e.ctx, elemPtr(points), elemPtr(scalars), elemPtr(points2), elemPtr(scalars2),
		C.size_t(n), elemPtr(out), bytePtr(status), fmtOf(c)))
	return
}

// BatchEncryptPerKey mirrors (*Ciphertext).Encrypt (elgamal/encrypt.go:42) with one public key per item.
func (e *Engine) BatchEncryptPerKey(pubKeys, k, m []fr.Element, c Coords) (ct []fr.Element, status []byte, err error) {
	n := len(k)
	ct, status = make([]fr.Element, 4*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_encrypt(e.ctx, elemPtr(pubKeys), 1, elemPtr(k), elemPtr(m), C.size_t(n), elemPtr(ct),
		bytePtr(status), fmtOf(c)))
	return
}

// CiphertextNeg mirrors (*Ciphertext).Neg (elgamal/ciphertext.go:37).
func (e *Engine) CiphertextNeg(a []fr.Element) (out []fr.Element, status []byte, err error) {
	n := len(a) / 4
	out, status = make([]fr.Element, 4*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_neg(e.ctx, elemPtr(a), C.size_t(n), elemPtr(out), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// CiphertextIsEqual mirrors (*Ciphertext).IsEqual (elgamal/ciphertext.go:79).
func (e *Engine) CiphertextIsEqual(a, b []fr.Element) (flags, status []byte, err error) {
	n := len(a) / 4
	flags, status = make([]byte, n), make([]byte, n)
	err = e.err(C.gcp_elgamal_is_equal(e.ctx, elemPtr(a), elemPtr(b), C.size_t(n), bytePtr(flags), bytePtr(status)))
	return
}

// CiphertextSelect mirrors (*Ciphertext).Select (elgamal/ciphertext.go:90): out[i] = sel[i] ? i1[i] : i2[i].
func (e *Engine) CiphertextSelect(sel []byte, i1, i2 []fr.Element) (out []fr.Element, status []byte, err error) {
	n := len(sel)
	out, status = make([]fr.Element, 4*n), make([]byte, n)
	err = e.err(C.gcp_elgamal_select(e.ctx, bytePtr(sel), elemPtr(i1), elemPtr(i2), C.size_t(n), elemPtr(out), bytePtr(status)))
	return
}

// AssertDecrypt mirrors (*Ciphertext).AssertDecrypt (elgamal/ciphertext.go:50): flags[i] = the gadget's assertions hold.
func (e *Engine) AssertDecrypt(ct, privKeys, msgs []fr.Element, c Coords) (flags, status []byte, err error) {
	n := len(privKeys)
	flags, status = make([]byte, n), make([]byte, n)
	err = e.err(C.gcp_elgamal_assert_decrypt(e.ctx, elemPtr(ct), elemPtr(privKeys), elemPtr(msgs), C.size_t(n), bytePtr(flags),
		bytePtr(status), fmtOf(c)))
	return
}

// VerifyDecryptionProofs mirrors DecryptionProof.Verify (elgamal/ciphertext.go:124) with hFn = poseidon.MultiHash.
func (e *Engine) VerifyDecryptionProofs(pubKeys, ct, msgs, a1, a2, z []fr.Element, c Coords) (flags, status []byte, err error) {
	n := len(z)
	flags, status = make([]byte, n), make([]byte, n)
	err = e.err(C.gcp_elgamal_verify_decryption_proof(e.ctx, elemPtr(pubKeys), elemPtr(ct), elemPtr(msgs), elemPtr(a1),
		elemPtr(a2), elemPtr(z), C.size_t(n), bytePtr(flags), bytePtr(status), fmtOf(c)))
	return
}

// EdDSAVerify mirrors eddsa.Verifier.IsValid (ecc/bn254/eddsa/verifier.go:55): A and R in iden3 (TE) coordinates.
func (e *Engine) EdDSAVerify(pubKeysTE, sigRTE, sigS, msgs []fr.Element) (flags, status []byte, err error) {
	n := len(sigS)
	flags, status = make([]byte, n), make([]byte, n)
	err = e.err(C.gcp_eddsa_verify(e.ctx, elemPtr(pubKeysTE), elemPtr(sigRTE), elemPtr(sigS), elemPtr(msgs), C.size_t(n),
		bytePtr(flags), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// FromTEtoRTE / FromRTEtoTE mirror ecc/format/twistededwards.go:29-48 over n points (X, Y pairs).  These are canonical-form
// helpers: the elements cross the boundary as canonical little-endian integers (fr.Element.Bytes reversed), not as
// fr.Element memory; a caller that holds fr.Elements uses the CoordsTE flag of the ElGamal calls instead.
func (e *Engine) FromTEtoRTE(pointsLE []byte) (out []byte, status []byte, err error) {
	n := len(pointsLE) / 64
	out, status = make([]byte, 64*n), make([]byte, n)
	err = e.err(C.gcp_te_to_rte(e.ctx, unsafe.Pointer(bytePtr(pointsLE)), C.size_t(n), unsafe.Pointer(bytePtr(out)), bytePtr(status)))
	return
}
func (e *Engine) FromRTEtoTE(pointsLE []byte) (out []byte, status []byte, err error) {
	n := len(pointsLE) / 64
	out, status = make([]byte, 64*n), make([]byte, n)
	err = e.err(C.gcp_rte_to_te(e.ctx, unsafe.Pointer(bytePtr(pointsLE)), C.size_t(n), unsafe.Pointer(bytePtr(out)), bytePtr(status)))
	return
}

// BatchMiMC7 mirrors mimc7 New / Write / Sum (hash/native/bn254/mimc7/mimc.go:47) over rows of `length` inputs (1..62).
func (e *Engine) BatchMiMC7(in []fr.Element, length int) (out []fr.Element, status []byte, err error) {
	if length <= 0 || len(in)%length != 0 {
		return nil, nil, errors.New("bad inputs provided")
	}
	n := len(in) / length
	out, status = make([]fr.Element, n), make([]byte, n)
	err = e.err(C.gcp_mimc7_hash(e.ctx, elemPtr(in), C.int(length), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// BallotBatch is the end-to-end voter batch (BASELINE config 5): per voter one census inclusion proof (arbo packed strings)
// and nFields encrypted values; the ciphertexts of the voters whose proof verifies are folded.  flags / status per voter,
// tally (4 * nFields elements) and tallyStatus per field.
func (e *Engine) BallotBatch(levels int, roots []fr.Element, packed [][]byte, keys, values []fr.Element, pubKey [2]fr.Element,
	k, m []fr.Element, nFields int, c Coords) (flags, status []byte, tally []fr.Element, tallyStatus []byte, err error) {
	n := len(keys)
	flags, status = make([]byte, n), make([]byte, n)
	tally, tallyStatus = make([]fr.Element, 4*nFields), make([]byte, nFields)
	blob, offsets := packBlob(packed)
	shared := 0
	if len(roots) == 1 && n != 1 {
		shared = 1
	}
	err = e.err(C.gcp_ballot_batch(e.ctx, C.int(levels), C.size_t(n), elemPtr(roots), C.int(shared), nil, bytePtr(blob),
		(*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(keys), elemPtr(values), unsafe.Pointer(&pubKey[0]), elemPtr(k),
		elemPtr(m), C.int(nFields), bytePtr(flags), bytePtr(status), elemPtr(tally), bytePtr(tallyStatus), fmtOf(c)))
	return
}

// PinnedElements allocates n field elements in page-locked host memory (gcp_host_alloc): host-buffer calls copy from
// it at full PCIe rate and overlap with the kernels.  Release with FreePinned(&s[0]).
func PinnedElements(n int) ([]fr.Element, error) {
	var p unsafe.Pointer
	if rc := C.gcp_host_alloc(C.size_t(n*32), &p); rc != 0 {
		return nil, fmt.Errorf("gcp_host_alloc failed (code %d)", int(rc))
	}
	return unsafe.Slice((*fr.Element)(p), n), nil
}

func FreePinned(first *fr.Element) { C.gcp_host_free(unsafe.Pointer(first)) }

// Group drives several GPUs of one box from this process (gcp_group_*): batches are sharded by index range, one host
// thread per device inside the C call; the tallies all-gather their partial ciphertexts with NCCL.
type Group struct{ grp *C.gcp_group }

func NewGroup(devices []int) (*Group, error) {
	d := make([]C.int, len(devices))
	for i, v := range devices {
		d[i] = C.int(v)
	}
	var g *C.gcp_group
	var p *C.int
	if len(d) > 0 {
		p = &d[0]
	}
	if rc := C.gcp_group_create(p, C.int(len(d)), nil, &g); rc != 0 {
		return nil, fmt.Errorf("gcp_b200: %s (code %d)", C.GoString(C.gcp_group_last_error(nil)), int(rc))
	}
	return &Group{grp: g}, nil
}

func (g *Group) Close()    { C.gcp_group_destroy(g.grp) }
func (g *Group) Size() int { return int(C.gcp_group_size(g.grp)) }

func (g *Group) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("gcp_b200: %s (code %d)", C.GoString(C.gcp_group_last_error(g.grp)), int(rc))
}

// BatchVerify is Engine.BatchVerify over all GPUs of the group.
func (g *Group) BatchVerify(p *Proofs) (flags, status []byte, err error) {
	n := len(p.Keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(p.Roots) == 1 && n != 1 {
		shared = 1
	}
	err = g.err(C.gcp_group_smt_verify(g.grp, C.int(p.Levels), C.size_t(n), elemPtr(p.Roots), C.int(shared),
		elemPtr(p.Siblings), elemPtr(p.OldKeys), elemPtr(p.OldValues), bytePtr(p.IsOld0), elemPtr(p.Keys), elemPtr(p.Values),
		bytePtr(p.Fnc), bytePtr(p.Enabled), bytePtr(flags), bytePtr(status), nil, C.GCP_FMT_MONTGOMERY))
	return
}

// Tally folds n_ballots x nFields ciphertexts over all GPUs (partial sums all-gathered with NCCL).
func (g *Group) Tally(ct []fr.Element, nFields int) (out []fr.Element, status []byte, err error) {
	nBallots := len(ct) / (4 * nFields)
	out, status = make([]fr.Element, 4*nFields), make([]byte, nFields)
	err = g.err(C.gcp_group_elgamal_tally(g.grp, elemPtr(ct), C.size_t(nBallots), C.int(nFields), elemPtr(out),
		bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// EncryptTally is the fused Encrypt + tally over all GPUs.
func (g *Group) EncryptTally(pubKey [2]fr.Element, k, m []fr.Element, nFields int) (out []fr.Element, status []byte, err error) {
	nBallots := len(k) / nFields
	out, status = make([]fr.Element, 4*nFields), make([]byte, nFields)
	err = g.err(C.gcp_group_elgamal_encrypt_tally(g.grp, unsafe.Pointer(&pubKey[0]), elemPtr(k), elemPtr(m),
		C.size_t(nBallots), C.int(nFields), elemPtr(out), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// BatchHash is Engine.BatchHash over all GPUs of the group.
func (g *Group) BatchHash(in []fr.Element, arity int) (out []fr.Element, status []byte, err error) {
	if arity <= 0 || len(in)%arity != 0 {
		return nil, nil, errors.New("bad inputs provided")
	}
	n := len(in) / arity
	out, status = make([]fr.Element, n), make([]byte, n)
	err = g.err(C.gcp_group_poseidon_hash(g.grp, elemPtr(in), C.int(arity), C.size_t(n), elemPtr(out), bytePtr(status),
		C.GCP_FMT_MONTGOMERY))
	return
}

// BatchEncrypt is Engine.BatchEncrypt (one shared key) over all GPUs of the group.
func (g *Group) BatchEncrypt(pubKey [2]fr.Element, k, m []fr.Element) (ct []fr.Element, status []byte, err error) {
	n := len(k)
	ct, status = make([]fr.Element, 4*n), make([]byte, n)
	err = g.err(C.gcp_group_elgamal_encrypt(g.grp, unsafe.Pointer(&pubKey[0]), 0, elemPtr(k), elemPtr(m), C.size_t(n),
		elemPtr(ct), bytePtr(status), C.GCP_FMT_MONTGOMERY))
	return
}

// BallotBatch is Engine.BallotBatch over all GPUs: voters sharded by index range, partial tallies all-gathered on the devices.
func (g *Group) BallotBatch(levels int, roots []fr.Element, packed [][]byte, keys, values []fr.Element, pubKey [2]fr.Element,
	k, m []fr.Element, nFields int) (flags, status []byte, tally []fr.Element, tallyStatus []byte, err error) {
	n := len(keys)
	flags, status = make([]byte, n), make([]byte, n)
	tally, tallyStatus = make([]fr.Element, 4*nFields), make([]byte, nFields)
	blob, offsets := packBlob(packed)
	shared := 0
	if len(roots) == 1 && n != 1 {
		shared = 1
	}
	err = g.err(C.gcp_group_ballot_batch(g.grp, C.int(levels), C.size_t(n), elemPtr(roots), C.int(shared), nil, bytePtr(blob),
		(*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(keys), elemPtr(values), unsafe.Pointer(&pubKey[0]), elemPtr(k),
		elemPtr(m), C.int(nFields), bytePtr(flags), bytePtr(status), elemPtr(tally), bytePtr(tallyStatus), C.GCP_FMT_MONTGOMERY))
	return
}

// UsesNCCL reports whether the group exchanges its partial tallies with ncclAllGather (more than one device).
func (g *Group) UsesNCCL() bool { return C.gcp_group_uses_nccl(g.grp) != 0 }

// BatchVerifyPacked is Engine.BatchVerifyPacked over all GPUs of the group.
func (g *Group) BatchVerifyPacked(p *Proofs, packed [][]byte) (flags, status []byte, err error) {
	n := len(p.Keys)
	flags, status = make([]byte, n), make([]byte, n)
	offsets := make([]uint64, n+1)
	total := 0
	for i, b := range packed {
		total += len(b)
		offsets[i+1] = uint64(total)
	}
	blob := make([]byte, 0, total+1)
	for _, b := range packed {
		blob = append(blob, b...)
	}
	blob = append(blob, 0)
	shared := 0
	if len(p.Roots) == 1 && n != 1 {
		shared = 1
	}
	err = g.err(C.gcp_group_smt_verify_packed(g.grp, C.int(p.Levels), C.size_t(n), elemPtr(p.Roots), C.int(shared),
		bytePtr(blob), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), elemPtr(p.OldKeys), elemPtr(p.OldValues),
		bytePtr(p.IsOld0), elemPtr(p.Keys), elemPtr(p.Values), bytePtr(p.Fnc), bytePtr(p.Enabled), bytePtr(flags),
		bytePtr(status), nil, C.GCP_FMT_MONTGOMERY))
	return
}

// PoseidonHint has the shape of a gnark solver.Hint (func(mod *big.Int, in, out []*big.Int) error; the reference
// registers one of that shape at hash/native/bn254/poseidon2/hints.go:10): out[0] = poseidon.Hash(in...) over BN254 Fr,
// computed by the engine.  For witness generation over many hashes prefer BatchHash: a hint call carries one hash.
func (e *Engine) PoseidonHint(_ *big.Int, in, out []*big.Int) error {
	if len(out) != 1 {
		return errors.New("PoseidonHint: one output expected")
	}
	elems := make([]fr.Element, len(in))
	for i, v := range in {
		elems[i].SetBigInt(v)
	}
	digest, status, err := e.BatchHash(elems, len(in))
	if err != nil {
		return err
	}
	if status[0] != StatusOK {
		return fmt.Errorf("PoseidonHint: status %d", status[0])
	}
	digest[0].BigInt(out[0])
	return nil
}

func setBig(e *fr.Element, v *big.Int) {
	if v != nil {
		e.SetBigInt(v)
	}
}

// FromAssignments flattens the witness records the reference's wrappers produce (smt.Assignment,
// tree/smt/wrapper.go:20-31; filled by WrapperArbo.Proof, wrapper_arbo.go:31-79) into the arrays BatchVerify takes:
// Fnc0 = 0 is an inclusion proof of (NewKey, NewValue), Fnc0 = 1 an exclusion proof of NewKey against
// (OldKey, OldValue, IsOld0) - the fnc input of smt.Verifier (verifier.go:102-111).
func FromAssignments(as []smt.Assignment) *Proofs {
	n := len(as)
	if n == 0 {
		return &Proofs{}
	}
	levels := len(as[0].Siblings)
	p := &Proofs{Levels: levels, Roots: make([]fr.Element, n), Siblings: make([]fr.Element, n*levels),
		OldKeys: make([]fr.Element, n), OldValues: make([]fr.Element, n), IsOld0: make([]byte, n),
		Keys: make([]fr.Element, n), Values: make([]fr.Element, n), Fnc: make([]byte, n)}
	for i := range as {
		a := &as[i]
		setBig(&p.Roots[i], a.OldRoot)
		for j, s := range a.Siblings {
			setBig(&p.Siblings[i*levels+j], s)
		}
		setBig(&p.OldKeys[i], a.OldKey)
		setBig(&p.OldValues[i], a.OldValue)
		setBig(&p.Keys[i], a.NewKey)
		setBig(&p.Values[i], a.NewValue)
		p.IsOld0[i], p.Fnc[i] = a.IsOld0, a.Fnc0
	}
	return p
}

// ProcessAssignments runs smt.Processor (tree/smt/processor.go:10) over state-transition records (WrapperArbo.Set /
// SetProof, wrapper_arbo.go:97-184) and returns the recomputed new roots; compare them with Assignment.NewRoot.
func (e *Engine) ProcessAssignments(as []smt.Assignment) (newRoots []fr.Element, status []byte, err error) {
	p := FromAssignments(as)
	fnc1 := make([]byte, len(as))
	for i := range as {
		fnc1[i] = as[i].Fnc1
	}
	return e.BatchProcess(p.Levels, p.Roots, p.Siblings, p.OldKeys, p.OldValues, p.IsOld0, p.Keys, p.Values, p.Fnc, fnc1)
}

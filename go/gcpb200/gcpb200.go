// NOTE: written against include/gcp_b200.h; never compiled (no Go toolchain in the build environment).
package gcpb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gnark_crypto_primitives_b200 -lgcp_b200 -Wl,-rpath,${SRCDIR}/../../gnark_crypto_primitives_b200
#include "gcp_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

type Engine struct{ ctx *C.gcp_ctx }

func New(device int) (*Engine, error) {
	var ctx *C.gcp_ctx
	if rc := C.gcp_ctx_create(C.int(device), nil, &ctx); rc != 0 {
		return nil, errors.New(C.GoString(C.gcp_last_error(nil)))
	}
	return &Engine{ctx}, nil
}
func (e *Engine) Close() { C.gcp_ctx_destroy(e.ctx) }

// BatchHash mirrors poseidon.Hash (hash/native/bn254/poseidon/poseidon.go:38) over n rows of `arity` inputs.
// []fr.Element memory is passed unchanged: GCP_FMT_MONTGOMERY is gnark-crypto's own representation.
func (e *Engine) BatchHash(in []fr.Element, arity int) ([]fr.Element, error) {
	n := len(in) / arity
	out := make([]fr.Element, n)
	status := make([]byte, n)
	rc := C.gcp_poseidon_hash(e.ctx, unsafe.Pointer(&in[0]), C.int(arity), C.size_t(n),
		unsafe.Pointer(&out[0]), (*C.uint8_t)(&status[0]), C.GCP_FMT_MONTGOMERY)
	if rc != 0 {
		return nil, errors.New(C.GoString(C.gcp_last_error(e.ctx))) // "bad inputs provided" for arity 0 or > 16
	}
	return out, nil
}

// BatchInclusionVerify mirrors smt.InclusionVerifier (tree/smt/verifier.go:29) over n Assignment records
// (tree/smt/wrapper.go:20-31) flattened by the caller: siblings is n*levels elements, root->leaf, zero padded.
func (e *Engine) BatchInclusionVerify(levels int, roots, siblings, keys, values []fr.Element) (flags, status []byte, err error) {
	n := len(keys)
	flags, status = make([]byte, n), make([]byte, n)
	shared := 0
	if len(roots) == 1 && n != 1 {
		shared = 1
	}
	rc := C.gcp_smt_verify_inclusion(e.ctx, C.int(levels), C.size_t(n), unsafe.Pointer(&roots[0]), C.int(shared),
		unsafe.Pointer(&siblings[0]), unsafe.Pointer(&keys[0]), unsafe.Pointer(&values[0]),
		(*C.uint8_t)(&flags[0]), (*C.uint8_t)(&status[0]), nil, C.GCP_FMT_MONTGOMERY)
	if rc != 0 {
		return nil, nil, errors.New(C.GoString(C.gcp_last_error(e.ctx)))
	}
	return flags, status, nil
}

"""ctypes binding of libgcp_b200.so (include/gcp_b200.h).  Fails loudly when the library is missing: the
package has no CPU path."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint8, c_uint64, c_void_p
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libgcp_b200.so"

GCP_OK = 0
GCP_ERR_BAD_ARG = -1
GCP_ERR_CUDA = -2
GCP_ERR_NO_DEVICE = -3
GCP_ERR_CONSTANTS = -4
GCP_ERR_ALLOC = -5

FMT_CANONICAL = 0
FMT_MONTGOMERY = 1
HASHER_POSEIDON = 0  # utils.PoseidonHasher (default)
HASHER_POSEIDON2 = 1  # utils.Poseidon2Hasher: the width-2 Merkle-Damgard hasher (gcp_ctx_set_smt_hasher)
MSG_U64 = 4         # or-ed into fmt of the fused tallies: messages are little-endian uint64 (8 bytes), not field elements
COORDS_TE = 2       # or-ed into fmt: curve points on the wire are in iden3 twisted-Edwards coordinates (gcp_b200.h)

STATUS_OK = 0
STATUS_NONCANONICAL = 1
STATUS_KEY_RANGE = 2
STATUS_NOT_BOOLEAN = 3
STATUS_OFF_CURVE = 4
STATUS_ZERO_DENOM = 5
STATUS_ASSERTION = 6
STATUS_MALFORMED = 7

_u8p = POINTER(c_uint8)

# name -> (restype, argtypes).  Must list every symbol include/gcp_b200.h declares (tests check this).
SIGNATURES = {
    "gcp_device_count": (c_int, []),
    "gcp_host_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "gcp_host_free": (None, [c_void_p]),
    "gcp_ctx_create": (c_int, [c_int, c_char_p, POINTER(c_void_p)]),
    "gcp_ctx_destroy": (None, [c_void_p]),
    "gcp_last_error": (c_char_p, [c_void_p]),
    "gcp_ctx_device": (c_int, [c_void_p]),
    "gcp_ctx_launch_count": (c_uint64, [c_void_p]),
    "gcp_probe_imad_wide": (c_int, [c_void_p, POINTER(ctypes.c_double)]),
    "gcp_poseidon_hash": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_poseidon_hash_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_poseidon_multihash": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_poseidon_multihash_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_mimc7_hash": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_mimc7_hash_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_poseidon2_set_round_keys": (c_int, [c_void_p, c_void_p, c_size_t, c_int]),
    "gcp_poseidon2_hash": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_poseidon2_hash_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_poseidon2_permutation": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_poseidon2_permutation_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_smt_verify": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_verify_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p]),
    "gcp_smt_verify_with_leaf_hash": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_verify_with_leaf_hash_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_smt_leaf_hash": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_smt_leaf_hash_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_smt_process_with_leaf_hash": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_process_with_leaf_hash_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_int, c_void_p]),
    "gcp_ctx_set_smt_hasher": (c_int, [c_void_p, c_int]),
    "gcp_ctx_smt_hasher": (c_int, [c_void_p]),
    "gcp_ctx_set_fixed_base_window": (c_int, [c_void_p, c_int]),
    "gcp_ctx_fixed_base_window": (c_int, [c_void_p, c_int]),
    "gcp_copy_threads": (c_int, []),
    "gcp_copy_probe": (c_int, [c_size_t, POINTER(ctypes.c_double)]),
    "gcp_smt_verify_packed": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int]),
    "gcp_smt_unpack_siblings_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                            c_int, c_void_p]),
    "gcp_group_create": (c_int, [POINTER(c_int), c_int, c_char_p, POINTER(c_void_p)]),
    "gcp_group_destroy": (None, [c_void_p]),
    "gcp_group_last_error": (c_char_p, [c_void_p]),
    "gcp_group_size": (c_int, [c_void_p]),
    "gcp_group_ctx": (c_void_p, [c_void_p, c_int]),
    "gcp_group_uses_nccl": (c_int, [c_void_p]),
    "gcp_group_poseidon_hash": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_group_smt_verify": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_group_smt_verify_packed": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_int]),
    "gcp_group_elgamal_encrypt": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                          c_int]),
    "gcp_group_elgamal_tally": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int]),
    "gcp_group_elgamal_encrypt_tally": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p,
                                                c_void_p, c_int]),
    "gcp_smt_scan_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gcp_smt_verify_inclusion": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_verify_exclusion": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_fixed_base_mul": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_fixed_base_mul_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_elgamal_scalar_mul": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                       c_int]),
    "gcp_elgamal_scalar_mul_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                           c_int, c_void_p]),
    "gcp_elgamal_encrypt": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_encrypt_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                        c_int, c_void_p]),
    "gcp_elgamal_add": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_add_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_elgamal_neg": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_tally": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_tally_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_elgamal_encrypt_tally": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p,
                                          c_int]),
    "gcp_elgamal_encrypt_tally_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p,
                                              c_void_p, c_int, c_void_p]),
    "gcp_keccak_address": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcp_keccak_address_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gcp_elgamal_is_equal": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gcp_elgamal_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gcp_ballot_batch": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int]),
    "gcp_group_ballot_batch": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int]),
    "gcp_ballot_batch_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_smt_process": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_process_packed": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_process_arbo": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gcp_smt_process_dev": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gcp_elgamal_assert_decrypt": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_elgamal_verify_decryption_proof": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                    c_size_t, c_void_p, c_void_p, c_int]),
    "gcp_te_to_rte": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gcp_rte_to_te": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gcp_eddsa_verify": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"gcp_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GCP_B200_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m gnark_crypto_primitives_b200.build` "
            "(this package has no CPU implementation)")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib

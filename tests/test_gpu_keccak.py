"""GPU parity: Ethereum address derivation (legacy Keccak-256) vs the oracle and public vectors."""
import random

import numpy as np
import pytest

from oracle import cport
from oracle import keccak as okeccak

pytestmark = pytest.mark.gpu


def test_address_of_secp256k1_generator(engine):
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    data = np.frombuffer(gx.to_bytes(32, "big") + gy.to_bytes(32, "big"), np.uint8)
    out = engine.keccak_address(data)
    assert out[0].tobytes().hex() == "7e5f4552091a69125d5dfcb7b8c2659029395bdf"  # public vector: private key 1


def test_random_batch_matches_oracle(engine):
    rng = np.random.default_rng(0xB200)
    n = 10007
    data = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    data[0] = 0
    data[1] = 255
    out = engine.keccak_address(data)
    want = cport.keccak_address(data, threads=8)
    assert (out == want).all()
    for i in (0, 1, 2, 500):
        assert out[i].tobytes() == okeccak.derive_address(data[i].tobytes())
    assert engine.keccak_address(np.zeros((0, 64), np.uint8)).shape == (0, 20)

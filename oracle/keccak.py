"""Oracle: legacy Keccak-256 and Ethereum address derivation.

Follows /root/reference/ecc/secp256k1/ecdsa/address.go:14-40: address = last 20 bytes of
legacy Keccak-256 (pad 0x01 .. 0x80, rate 136) over X_be32 || Y_be32, packed big-endian into one
field variable (utils/uints.go:33-48).  Keccak-f[1600] itself lives in gnark std/hash/sha3
(un-vendored); restated from the published Keccak specification and pinned by public vectors
(Keccak-256(""), Keccak-256("hello"), address of secp256k1 G) in tests/test_oracle_golden.py.
"""
RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M64 if n else x


def keccak_f1600(a):
    """a: 25 lanes, index x + 5*y."""
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= RC[rnd]
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            st[i] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        st = keccak_f1600(st)
    return b"".join(x.to_bytes(8, "little") for x in st[:4])


def derive_address(pub_xy_be: bytes) -> bytes:
    """address.go:14-40 on a 64-byte X_be || Y_be public key -> 20 address bytes."""
    assert len(pub_xy_be) == 64
    return keccak256(pub_xy_be)[12:]

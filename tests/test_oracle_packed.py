"""CPU: the arbo packed-sibling restatement (oracle/smt.py) round-trips, matches the verifier's input shape on an
oracle-built tree, and rejects what arbo.UnpackSiblings rejects.  No vector for this format exists in the reference
tree (arbo is an un-vendored dependency): parity for the wire format itself is UNPINNED, see DESIGN.md section 7."""
import random

from oracle import smt as osmt


def test_pack_layout_is_the_documented_one():
    sib = [0, 5, 0, 0, 7, 0, 0, 0, 9]
    b = osmt.pack_siblings(sib)
    assert int.from_bytes(b[0:2], "little") == len(b) == 4 + 2 + 3 * 32
    assert int.from_bytes(b[2:4], "little") == 2
    assert b[4] == 0b00010010 and b[5] == 0b00000001
    assert b[6:38] == (5).to_bytes(32, "little") and b[70:102] == (9).to_bytes(32, "little")
    assert osmt.unpack_siblings(b) == sib


def test_round_trip_and_padding():
    rng = random.Random(7)
    for _ in range(200):
        depth = rng.randint(0, 40)
        sib = [0 if rng.random() < 0.3 else rng.getrandbits(253) for _ in range(depth)]
        b = osmt.pack_siblings(sib)
        un = osmt.unpack_siblings(b)
        # trailing zero siblings are not materialised (arbo stops when the data is exhausted); padding restores them
        assert un == sib[:len(un)] and all(s == 0 for s in sib[len(un):])
        padded, st = osmt.assignment_siblings(b, 64)
        assert st == 0 and padded == (sib + [0] * 64)[:64]
        cut, st = osmt.assignment_siblings(b, 8)
        assert st == 0 and cut == (sib + [0] * 8)[:8]


def test_tree_proofs_verify_from_their_packed_form():
    rng = random.Random(11)
    tree = osmt.Tree(64)
    keys = [rng.getrandbits(64) for _ in range(20)]
    for k in keys:
        tree.add(k, rng.getrandbits(200))
    root = tree.root()
    for k in keys[:6]:
        p = tree.gen_proof(k)
        sib, st = osmt.assignment_siblings(tree.last_packed, 64)
        assert st == 0 and sib == p["siblings"]
        assert osmt.inclusion_verifier(root, sib, k, p["old_value"])[0] == 1


def test_malformed_strings_are_rejected():
    good = osmt.pack_siblings([3, 0, 4])
    assert osmt.unpack_siblings(good[:-1]) is None            # length field != len
    assert osmt.unpack_siblings(good + b"\x00") is None
    assert osmt.unpack_siblings(b"\x03\x00\x00") is None      # shorter than the header
    bad_l = (6).to_bytes(2, "little") + (9).to_bytes(2, "little") + b"\x00\x00"
    assert osmt.unpack_siblings(bad_l) is None                # bitmap runs past the string
    cut = (4 + 1 + 40).to_bytes(2, "little") + (1).to_bytes(2, "little") + b"\x03" + bytes(40)
    assert osmt.unpack_siblings(cut) is None                  # second set bit has only 8 of its 32 bytes
    assert osmt.assignment_siblings(cut, 16) == ([0] * 16, osmt.STATUS_MALFORMED)
    # more set bits than whole siblings: arbo stops at the end of the data, no error
    short = (4 + 1 + 32).to_bytes(2, "little") + (1).to_bytes(2, "little") + b"\x07" + (9).to_bytes(32, "little")
    assert osmt.unpack_siblings(short) == [9]
    assert osmt.unpack_siblings(osmt.pack_siblings([])) == []

"""Oracle: circomlib/arbo sparse-Merkle-tree proof verifier (value semantics of the gadget).

Follows /root/reference/tree/smt:
  verifier.go:29-43 InclusionVerifier, :66-81 ExclusionVerifier, :102-121 Verifier,
  :171-242 VerifierWithLeafHashFlag
  lev_ins.go:43-77 LevInsFlag, verifier_sm.go:5-14 VerifierSM,
  verifier_level.go:8-17 VerifierLevel, hash.go:10-27 Hash1/Hash2,
  utils.go:11-56 lowBits / IsEqual / ForceEqualIfEnabledFlag / MultiAnd / Switcher
Every api.Add/Sub/Mul is evaluated mod r exactly as gnark's test engine does.

Two failure classes:
  flag   -- the gadget's 0/1 output;
  status -- non-zero when the reference would fail an *assertion* (solver error):
            STATUS_NONCANONICAL  an input >= r
            STATUS_KEY_RANGE     key >= 2^n            (utils.go:11-13, bits.ToBinary)
            STATUS_NOT_BOOLEAN   enabled / fnc / isOld0 not in {0,1} (api.Select / api.And assert booleans)

`Tree` is a minimal in-memory arbo-shaped tree (leaf = H(k,v,1), node = H(l,r), empty = 0,
path bit i = bit i of the key, leaf at the shallowest depth where its path is unique) used to
produce real inclusion/exclusion proofs with the fixture shape of testutil/utils.go:95-185 and
wrapper_arbo.go:31-79 (siblings root->leaf, zero padded to `levels`).
"""
from . import poseidon
from .field import R

STATUS_OK = 0
STATUS_NONCANONICAL = 1
STATUS_KEY_RANGE = 2
STATUS_NOT_BOOLEAN = 3


def hash1(key, *values):
    """hash.go:10-19: H(key, values..., 1); one value per leaf in smt.Verifier / smt.Processor."""
    return poseidon.hash([key, *values, 1])


def hash2(l, r):
    """hash.go:21-27."""
    return poseidon.hash([l, r])


def lev_ins_flag(enabled, siblings):
    """lev_ins.go:43-77 -> (valid, levIns[])."""
    n = len(siblings)
    lev_ins = [0] * n
    if n < 2:
        return (0 if enabled else 1), lev_ins
    is_zero = [1 if s % R == 0 else 0 for s in siblings]
    done = [0] * (n - 1)
    lev_ins[n - 1] = (1 - is_zero[n - 2]) % R
    done[n - 2] = lev_ins[n - 1]
    for i in range(n - 2, 0, -1):
        lev_ins[i] = (1 - done[i]) * (1 - is_zero[i - 1]) % R
        done[i - 1] = (lev_ins[i] + done[i]) % R
    lev_ins[0] = (1 - done[0]) % R
    leaf_zero_ok = is_zero[n - 1]
    one_hot = 1 if (sum(lev_ins) - 1) % R == 0 else 0
    valid = (leaf_zero_ok & one_hot) if enabled else 1
    return valid, lev_ins


def verifier_sm(is0, lev_ins, fnc, prev_top, prev_i0, prev_iold, prev_inew, prev_na):
    """verifier_sm.go:5-14."""
    aux1 = prev_top * lev_ins % R
    aux2 = aux1 * fnc % R
    st_top = (prev_top - aux1) % R
    st_inew = (aux1 - aux2) % R
    st_iold = aux2 * ((1 - is0) % R) % R
    st_i0 = aux1 * is0 % R
    st_na = (prev_na + prev_inew + prev_iold + prev_i0) % R
    return st_top, st_i0, st_iold, st_inew, st_na


def verifier_with_leaf_hash_flag(enabled, root, siblings, old_key, hash1_old, is_old0, key, hash1_new, fnc,
                                 count_hashes=None):
    """verifier.go:171-242 -> (flag, status, computed_root).

    All n Hash2 calls are evaluated, masked levels included, exactly like the gadget.
    """
    n = len(siblings)
    vals = [enabled, root, old_key, hash1_old, is_old0, key, hash1_new, fnc] + list(siblings)
    if any(not (0 <= int(v) < R) for v in vals):
        return 0, STATUS_NONCANONICAL, 0
    if any(b not in (0, 1) for b in (enabled, fnc, is_old0)):
        return 0, STATUS_NOT_BOOLEAN, 0
    if key >> n:
        return 0, STATUS_KEY_RANGE, 0
    n2b_new = [(key >> i) & 1 for i in range(n)]                     # utils.go:11-13, LSB first
    flag_lev_ins, lev_ins = lev_ins_flag(enabled, siblings)

    st = []
    prev = (enabled, 0, 0, 0, (1 - enabled) % R)                      # verifier.go:196-197
    for i in range(n):
        prev = verifier_sm(is_old0, lev_ins[i], fnc, *prev)
        st.append(prev)
    st_top = [s[0] for s in st]
    st_i0 = [s[1] for s in st]
    st_iold = [s[2] for s in st]
    st_inew = [s[3] for s in st]
    st_na = [s[4] for s in st]
    sum_states = (st_na[-1] + st_iold[-1] + st_inew[-1] + st_i0[-1]) % R
    flag_states = 1 if sum_states == 1 else 0

    levels = [0] * n
    for i in range(n - 1, -1, -1):                                    # verifier.go:211-220
        child = levels[i + 1] if i < n - 1 else 0
        if n2b_new[i]:
            l, r_ = siblings[i], child                                # Switcher utils.go:50-56
        else:
            l, r_ = child, siblings[i]
        proof_hash = hash2(l, r_)
        if count_hashes is not None:
            count_hashes[0] += 1
        levels[i] = (proof_hash * st_top[i] + hash1_old * st_iold[i] + hash1_new * st_inew[i]) % R

    are_keys_equal = 1 if old_key == key else 0
    key_reuse = fnc & ((1 - is_old0) % R) & are_keys_equal & enabled
    flag_key_reuse = 1 if key_reuse == 0 else 0
    flag_root = (1 if levels[0] == root else 0) if enabled else 1
    flag = flag_states & flag_key_reuse & flag_root & flag_lev_ins
    return flag, STATUS_OK, levels[0]


def verifier(enabled, root, siblings, old_key, old_value, is_old0, key, value, fnc):
    """verifier.go:102-121 -> (flag, status, computed_root)."""
    for v in (old_key, old_value, key, value):
        if not (0 <= int(v) < R):
            return 0, STATUS_NONCANONICAL, 0
    h_old = hash1(old_key, old_value)
    h_new = hash1(key, value)
    return verifier_with_leaf_hash_flag(enabled, root, siblings, old_key, h_old, is_old0, key, h_new, fnc)


def inclusion_verifier(root, siblings, key, value):
    """verifier.go:29-43."""
    return verifier(1, root, siblings, key, value, 0, key, value, 0)


def exclusion_verifier(root, siblings, old_key, old_value, is_old0, key):
    """verifier.go:66-81."""
    return verifier(1, root, siblings, old_key, old_value, is_old0, key, 0, 1)


def fold_inclusion(siblings, key, value):
    """The reduction SURVEY.md 8a4 derives for fnc=0, enabled=1 (used to synthesise roots quickly)."""
    n = len(siblings)
    last = -1
    for i in range(n - 1):
        if siblings[i] != 0:
            last = i
    lidx = last + 1
    acc = hash1(key, value)
    for i in range(lidx - 1, -1, -1):
        acc = hash2(siblings[i], acc) if (key >> i) & 1 else hash2(acc, siblings[i])
    return acc


# --------------------------------------------------------------------------------------
# arbo-shaped in-memory tree (fixture generator)
# --------------------------------------------------------------------------------------
class Tree:
    """Sparse Merkle tree with arbo/circomlib shape.  Nodes: None (empty), ('leaf', k, v), ('node', l, r)."""

    def __init__(self, max_levels):
        self.max_levels = max_levels
        self.root_node = None
        self._hash_cache = {}

    def _h(self, node):
        if node is None:
            return 0
        hid = id(node)
        got = self._hash_cache.get(hid)
        if got is not None and got[0] is node:
            return got[1]
        if node[0] == "leaf":                                  # a tuple value is a multi-value leaf: Hash1(key, values...)
            h = hash1(node[1], *node[2]) if isinstance(node[2], tuple) else hash1(node[1], node[2])
        else:
            h = hash2(self._h(node[1]), self._h(node[2]))
        self._hash_cache[hid] = (node, h)
        return h

    def root(self):
        return self._h(self.root_node)

    def add(self, key, value):
        assert 0 <= key < (1 << self.max_levels)
        self.root_node = self._add(self.root_node, key, value, 0)

    def _add(self, node, key, value, depth):
        if node is None:
            return ("leaf", key, value)
        if node[0] == "leaf":
            if node[1] == key:
                return ("leaf", key, value)
            return self._split(node, ("leaf", key, value), depth)
        if (key >> depth) & 1:
            return ("node", node[1], self._add(node[2], key, value, depth + 1))
        return ("node", self._add(node[1], key, value, depth + 1), node[2])

    def _split(self, a, b, depth):
        if depth >= self.max_levels:
            raise ValueError("max levels reached")
        ba, bb = (a[1] >> depth) & 1, (b[1] >> depth) & 1
        if ba != bb:
            return ("node", b, a) if ba else ("node", a, b)
        sub = self._split(a, b, depth + 1)
        return ("node", None, sub) if ba else ("node", sub, None)

    def gen_proof(self, key):
        """-> dict(exists, old_key, old_value, is_old0, siblings) ; siblings root->leaf, zero padded."""
        sibs = []
        node = self.root_node
        depth = 0
        while node is not None and node[0] == "node":
            if (key >> depth) & 1:
                sibs.append(self._h(node[1]))
                node = node[2]
            else:
                sibs.append(self._h(node[2]))
                node = node[1]
            depth += 1
        packed = pack_siblings(sibs)
        sibs += [0] * (self.max_levels - len(sibs))
        self.last_packed = packed  # arbo GenProof's siblingsPacked for the proof just generated
        if node is None:
            return dict(exists=False, old_key=0, old_value=0, is_old0=1, siblings=sibs)
        if node[1] == key:
            return dict(exists=True, old_key=key, old_value=node[2], is_old0=0, siblings=sibs)
        return dict(exists=False, old_key=node[1], old_value=node[2], is_old0=0, siblings=sibs)


# --------------------------------------------------------------------------------------
# arbo packed-sibling wire format (what GenProof returns and the callers unpack)
# --------------------------------------------------------------------------------------
STATUS_MALFORMED = 7  # arbo.UnpackSiblings would return an error (or slice out of range) on this byte string


def pack_siblings(siblings, hash_len=32):
    """arbo.PackSiblings (github.com/vocdoni/arbo v0.0.0-20250707215550-6dee1243bb29, un-vendored; the reference
    only holds the call sites of its inverse: tree/smt/wrapper_arbo.go:64,166 and testutil/utils.go:153).
    Layout:  [2 B full length, LE][2 B bitmap length L, LE][L B bitmap][32 B per non-zero sibling, LE integers]
    bitmap bit i (byte i/8, LSB first) is set iff sibling i is non-zero.  `siblings` is the root->leaf list as
    GenProof produces it (length = depth of the leaf, not padded).  PARITY UNPINNED: restated from the published
    arbo source, no vector for it exists in the reference tree."""
    bitmap = bytearray((len(siblings) + 7) // 8)
    body = b""
    for i, sib in enumerate(siblings):
        if sib != 0:
            bitmap[i // 8] |= 1 << (i % 8)
            body += int(sib).to_bytes(hash_len, "little")
    full = 4 + len(bitmap) + len(body)
    if full > 0xFFFF or len(bitmap) > 0xFFFF:
        raise ValueError("packed siblings too long")
    return full.to_bytes(2, "little") + len(bitmap).to_bytes(2, "little") + bytes(bitmap) + body


def unpack_siblings(b, hash_len=32):
    """arbo.UnpackSiblings: -> list of integers (root->leaf), or None where arbo returns an error / the Go slice
    expression would run past the data (treated alike: STATUS_MALFORMED).  Trailing clear bits after the data is
    exhausted are not materialised (the caller pads with zeros anyway, wrapper_arbo.go:69-76)."""
    if len(b) < 4:
        return None
    full = int.from_bytes(b[0:2], "little")
    l = int.from_bytes(b[2:4], "little")
    if full != len(b) or 4 + l > len(b):
        return None
    bitmap = b[4:4 + l]
    data = b[4 + l:]
    out = []
    pos = 0
    for i in range(8 * l):
        if pos >= len(data):
            break
        if (bitmap[i // 8] >> (i % 8)) & 1:
            if pos + hash_len > len(data):
                return None
            out.append(int.from_bytes(data[pos:pos + hash_len], "little"))
            pos += hash_len
        else:
            out.append(0)
    return out


def assignment_siblings(packed, levels):
    """Assignment.Siblings as wrapper_arbo.go:63-76 / testutil/utils.go:152-166 build it: unpack, then pad with
    zeros (or silently cut) to `levels`.  -> (siblings, status)."""
    un = unpack_siblings(packed)
    if un is None:
        return [0] * levels, STATUS_MALFORMED
    return [un[i] if i < len(un) else 0 for i in range(levels)], STATUS_OK


def arbo_add_or_update(tree, key, value):
    """WrapperArbo.SetWithTx / addOrUpdate (tree/smt/wrapper_arbo.go:97-185) on the oracle's Tree: the Assignment the
    reference's callers feed to smt.Processor, with its siblings taken from a proof generated AFTER the change.
    -> dict(old_root, new_root, old_key, old_value, is_old0, new_key, new_value, fnc0, fnc1, siblings, packed, status)
    `packed` is GenProof's siblingsPacked (post-insert); `siblings` what :166-179 make of it."""
    old_root = tree.root()                                            # :125-129
    found = tree.gen_proof(key)                                       # GetWithTx :135: the leaf on the key's path, if any
    fnc0, fnc1 = (0, 1) if found["exists"] else (1, 0)                # update :113-117 / add :107-111
    tree.add(key, value)
    is_old0 = 1 if found["is_old0"] else 0                            # :146-150 (len(oldKeyBytes) > 0)
    new_root = tree.root()                                            # :152-156
    tree.gen_proof(key)                                               # :158
    packed = tree.last_packed
    un = unpack_siblings(packed)                                      # :166
    status = STATUS_OK
    if un is None:
        un, status = [], STATUS_MALFORMED
    elif is_old0 == 0 and fnc1 == 0:                                  # :170-172
        if not un:
            status = STATUS_MALFORMED                                 # the Go slice expression [0:-1] panics
        un = un[:-1]
    sib = [un[i] if i < len(un) else 0 for i in range(tree.max_levels)]   # :174-181
    return dict(old_root=old_root, new_root=new_root, old_key=found["old_key"], old_value=found["old_value"],
                is_old0=is_old0, new_key=key, new_value=value, fnc0=fnc0, fnc1=fnc1, siblings=sib, packed=packed,
                status=status)


# --------------------------------------------------------------------------------------
# Processor (state transition: insert / update / delete / nop)
# --------------------------------------------------------------------------------------
STATUS_ASSERTION = 6  # an AssertIsEqual of the processor gadget fails (old root mismatch, LevIns, states, key rule)


def processor_sm(xor, is0, lev_ins, fnc0, prev_top, prev_old0, prev_bot, prev_new1, prev_na, prev_upd):
    """processor_sm.go:7-17."""
    aux1 = prev_top * lev_ins % R
    aux2 = aux1 * fnc0 % R
    st_top = (prev_top - aux1) % R
    st_old0 = aux2 * is0 % R
    st_new1 = (aux2 - st_old0 + prev_bot) * xor % R
    st_bot = (1 - xor) * ((aux2 - st_old0 + prev_bot) % R) % R
    st_upd = (aux1 - aux2) % R
    st_na = (prev_new1 + prev_old0 + prev_na + prev_upd) % R
    return st_top, st_old0, st_bot, st_new1, st_na, st_upd


def processor_level(st_top, st_old0, st_bot, st_new1, st_upd, sibling, old1leaf, new1leaf, newlrbit, old_child, new_child):
    """processor_level.go:10-27 (both Hash2 calls evaluated, as in the gadget)."""
    l, r_ = (sibling, old_child) if newlrbit else (old_child, sibling)            # Switcher
    old_proof_hash = hash2(l, r_)
    old_root = (old1leaf * ((st_bot + st_new1 + st_upd) % R) + old_proof_hash * st_top) % R
    a = (new_child * ((st_top + st_bot) % R) + new1leaf * st_new1) % R
    b = (sibling * st_top + old1leaf * st_new1) % R
    l, r_ = (b, a) if newlrbit else (a, b)
    new_proof_hash = hash2(l, r_)
    new_root = (new_proof_hash * ((st_top + st_bot + st_new1) % R) + new1leaf * ((st_old0 + st_upd) % R)) % R
    return old_root, new_root


def processor(old_root, siblings, old_key, old_value, is_old0, new_key, new_value, fnc0, fnc1):
    """processor.go:10-14 -> (new_root, status).  fnc = (1,0) insert, (0,1) update, (1,1) delete, (0,0) nop."""
    for v in (old_key, old_value, new_key, new_value):
        if not (0 <= int(v) < R):
            return 0, STATUS_NONCANONICAL
    return processor_with_leaf_hash(old_root, siblings, old_key, hash1(old_key, old_value), is_old0, new_key,
                                    hash1(new_key, new_value), fnc0, fnc1)


def processor_with_leaf_hash(old_root, siblings, old_key, hash1_old, is_old0, new_key, hash1_new, fnc0, fnc1):
    """processor.go:16-72 -> (new_root, status)."""
    n = len(siblings)
    vals = [old_root, old_key, hash1_old, new_key, hash1_new] + list(siblings)
    if any(not (0 <= int(v) < R) for v in vals):
        return 0, STATUS_NONCANONICAL
    if any(b not in (0, 1) for b in (is_old0, fnc0, fnc1)):
        return 0, STATUS_NOT_BOOLEAN                                   # AssertIsBoolean / api.Select / api.And
    if (old_key >> n) or (new_key >> n):
        return 0, STATUS_KEY_RANGE                                     # lowBits, processor.go:23-24
    enabled = (fnc0 + fnc1 - fnc0 * fnc1) % R
    valid, lev_ins = lev_ins_flag(enabled, siblings)                   # LevIns asserts valid == 1 (lev_ins.go:16-20)
    if valid != 1:
        return 0, STATUS_ASSERTION
    xors = [((old_key >> i) & 1) ^ ((new_key >> i) & 1) for i in range(n)]
    st = []
    prev = (enabled, 0, 0, 0, (1 - enabled) % R, 0)
    for i in range(n):
        prev = processor_sm(xors[i], is_old0, lev_ins[i], fnc0, *prev)
        st.append(prev)
    top, old0, bot, new1, na, upd = st[-1]
    if (na + new1 + old0 + upd) % R != 1:                              # processor.go:47
        return 0, STATUS_ASSERTION
    old_lv, new_lv = [0] * n, [0] * n
    for i in range(n - 1, -1, -1):
        oc = old_lv[i + 1] if i < n - 1 else 0
        nc = new_lv[i + 1] if i < n - 1 else 0
        t, o0, b, n1, _, u = st[i]
        old_lv[i], new_lv[i] = processor_level(t, o0, b, n1, u, siblings[i], hash1_old, hash1_new,
                                               (new_key >> i) & 1, oc, nc)
    both = fnc0 * fnc1
    top_l, top_r = (new_lv[0], old_lv[0]) if both else (old_lv[0], new_lv[0])   # Switcher(fnc0*fnc1, old, new)
    if enabled and old_root != top_l:                                  # ForceEqualIfEnabled, processor.go:60
        return 0, STATUS_ASSERTION
    new_root = (enabled * ((top_r - old_root) % R) + old_root) % R
    keys_equal = 1 if old_key == new_key else 0
    if (1 - fnc0) * fnc1 * (1 - keys_equal) != 0:                      # processor.go:64-70
        return 0, STATUS_ASSERTION
    return new_root, STATUS_OK

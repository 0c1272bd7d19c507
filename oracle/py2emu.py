"""TEST INFRASTRUCTURE ONLY (oracle/): CPython-2.7 container semantics that leak into the
reference's output bytes.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.

The reference is Python 2 code whose row orders come from dict / set iteration order
(reference falcon_unzip/phasing.py:466; rr_hctg_track.py:99,113,120,126).  These
emulators restate CPython 2.7's Objects/dictobject.c + setobject.c insertion /
resize / probe rules (no deletions occur on the path) and Objects/stringobject.c
string_hash with hash randomisation off (the 2.7 default).

Pinned by hand-derived vectors in tests/test_py2emu.py;
no Python 2 interpreter exists in the build container, so these are otherwise unpinned.
"""
from __future__ import annotations

from typing import Callable, Hashable, Iterable, List

_M64 = (1 << 64) - 1


def py27_str_hash(s: str) -> int:
    """Objects/stringobject.c:string_hash, 64-bit long, no randomisation."""
    if not s:
        return 0
    b = s.encode("latin-1")
    x = (b[0] << 7) & _M64
    for c in b:
        x = ((1000003 * x) & _M64) ^ c
    x ^= len(b)
    if x >= 1 << 63:
        x -= 1 << 64
    return -2 if x == -1 else x


def py27_int_hash(i: int) -> int:
    return -2 if i == -1 else i


class _Py27Table:
    """Open-addressing table of dictobject.c / setobject.c (insert-only)."""

    def __init__(self):
        self.mask = 7
        self.slots: List = [None] * 8     # (hash, key) or None
        self.used = 0

    def _find(self, slots, mask, h, key):
        i = h & mask                      # (size_t)hash & mask
        perturb = h & _M64
        while True:
            e = slots[i & mask]
            if e is None or (key is not None and e[0] == h and e[1] == key):
                return i & mask
            i = (i * 5 + perturb + 1) & _M64
            perturb >>= 5

    def insert(self, h: int, key) -> bool:
        s = self._find(self.slots, self.mask, h, key)
        if self.slots[s] is not None:
            return False
        self.slots[s] = (h, key)
        self.used += 1
        if self.used * 3 >= (self.mask + 1) * 2:      # fill == used (no dummies)
            self._resize((2 if self.used > 50000 else 4) * self.used)
        return True

    def _resize(self, minused: int):
        newsize = 8
        while newsize <= minused:
            newsize <<= 1
        new = [None] * newsize
        for e in self.slots:
            if e is not None:
                new[self._find(new, newsize - 1, e[0], None)] = e
        self.slots, self.mask = new, newsize - 1

    def keys(self):
        return [e[1] for e in self.slots if e is not None]


def py27_order(keys_in_insertion_order: Iterable[Hashable],
               hashfn: Callable[[Hashable], int]) -> List:
    t = _Py27Table()
    for k in keys_in_insertion_order:
        t.insert(hashfn(k), k)
    return t.keys()


def py27_int_dict_order(keys: Iterable[int]) -> List[int]:
    """Iteration order of a py2 dict whose int keys were inserted in this order."""
    return py27_order(keys, py27_int_hash)


def py27_str_dict_order(keys: Iterable[str]) -> List[str]:
    return py27_order(keys, py27_str_hash)


class Py27StrSet:
    """Stand-in for a py2 ``set`` of str: remembers insertion order, iterates in py2 order."""

    def __init__(self, items: Iterable[str] = ()):
        self._d = {}
        for x in items:
            self.add(x)

    def add(self, x: str) -> None:
        self._d[x] = None

    def __contains__(self, x) -> bool:
        return x in self._d

    def __len__(self) -> int:
        return len(self._d)

    def __iter__(self):
        return iter(py27_str_dict_order(self._d))


def py27_float_str(v: float) -> str:
    """Python 2 ``str(float)``: '%.12g', plus '.0' when the result looks integral."""
    s = "%.12g" % v
    if "." not in s and "e" not in s and "n" not in s:      # n: inf / nan
        s += ".0"
    return s


class Py27Float(float):
    """A float whose str() / print form is Python 2's."""

    def __str__(self):
        return py27_float_str(float(self))

    __repr__ = __str__


def py27_tuple_hash(t) -> int:
    """Objects/tupleobject.c:tuplehash for tuples of int / str items."""
    x, mult, n = 0x345678, 1000003, len(t)
    for it in t:
        n -= 1
        h = py27_str_hash(it) if isinstance(it, str) else py27_int_hash(it)
        x = ((x ^ (h & _M64)) * mult) & _M64
        mult = (mult + 82520 + n + n) & _M64
    x = (x + 97531) & _M64
    if x >= 1 << 63:
        x -= 1 << 64
    return -2 if x == -1 else x


def py27_tuple_dict_order(keys):
    return py27_order(keys, py27_tuple_hash)

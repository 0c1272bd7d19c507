"""CPython-2.7 container semantics that leak into the reference's output bytes (SURVEY.md
Appendix B): str hash (Objects/stringobject.c, no hash randomisation) and the insert-only
open-addressing table shared by dict and set (Objects/dictobject.c, setobject.c).  Host-side
formatting only; used for the row orders of rawread_to_contigs (B.4)."""
from __future__ import annotations

from typing import Iterable, List

_M64 = (1 << 64) - 1


def str_hash(s: str) -> int:
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


def table_order(keys: Iterable, hashes: Iterable[int]) -> List:
    """Iteration order (slot order) of a py2 dict / set after inserting `keys` in order."""
    mask, slots, used = 7, [None] * 8, 0
    for key, h in zip(keys, hashes):
        i = h & mask
        perturb = h & _M64
        while True:
            e = slots[i & mask]
            if e is None:
                break
            if e[0] == h and e[1] == key:
                i = -1
                break
            i = (i * 5 + perturb + 1) & _M64
            perturb >>= 5
        if i == -1:
            continue
        slots[i & mask] = (h, key)
        used += 1
        if used * 3 >= (mask + 1) * 2:
            newsize = 8
            minused = (2 if used > 50000 else 4) * used
            while newsize <= minused:
                newsize <<= 1
            new = [None] * newsize
            nm = newsize - 1
            for e in slots:
                if e is None:
                    continue
                j = e[0] & nm
                p = e[0] & _M64
                while new[j & nm] is not None:
                    j = (j * 5 + p + 1) & _M64
                    p >>= 5
                new[j & nm] = e
            slots, mask = new, nm
    return [e[1] for e in slots if e is not None]


def str_dict_order(keys: Iterable[str]) -> List[str]:
    keys = list(keys)
    return table_order(keys, [str_hash(k) for k in keys])


def int_hash(i: int) -> int:
    return -2 if i == -1 else i


def tuple_hash(item_hashes) -> int:
    """Objects/tupleobject.c:tuplehash (CPython 2.7, 64-bit long) from the hashes of the items."""
    x, mult, n = 0x345678, 1000003, len(item_hashes)
    for h in item_hashes:
        n -= 1
        x = ((x ^ (h & _M64)) * mult) & _M64
        mult = (mult + 82520 + n + n) & _M64
    x = (x + 97531) & _M64
    if x >= 1 << 63:
        x -= 1 << 64
    return -2 if x == -1 else x

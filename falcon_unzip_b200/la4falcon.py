"""LA4Falcon text parsed on the device (fuz_parse_la4falcon): the concatenated output of `LA4Falcon -m / -mo`
for one or more LAS files goes to the GPU once; the column arrays stay there for fuz_rr_track /
fuz_ovlp_filter.  Host copies of the columns are made only when somebody asks for them (formatting of
selected lines, tie groups).  PyTorch owns the buffers; there is no CPU fallback for the parse itself -- only
the identity column of lines the kernel flags (exponent / inf / nan notation, more than 15 significant digits)
is evaluated with Python's float(), as the reference does for every line."""
from __future__ import annotations

import bisect
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, engine
from ._lib import FuzError, lib

COLS = ("q", "t", "len", "qs", "qe", "ql", "ts", "te", "tl")
_PY2_FLOAT = re.compile(r"^[+-]?((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?|inf|infinity|nan)$", re.I)


def normalise_blobs(blobs: Sequence) -> List[bytes]:
    out = []
    for b in blobs:
        if not isinstance(b, (bytes, bytearray, memoryview)):        # an iterable of text lines
            b = "".join(x if x.endswith("\n") else x + "\n" for x in b).encode("ascii")
        b = bytes(b)
        out.append(b if not b or b.endswith(b"\n") else b + b"\n")
    return out


class DeviceLines:
    """Columns of LA4Falcon lines as device tensors (`d`), the text on the host (`text`)."""

    def __init__(self, blobs: Sequence, require_id9: bool):
        import torch
        import warnings
        self.blobs = normalise_blobs(blobs)
        self._text: Optional[bytes] = None
        eng = engine.get_engine()
        dev = eng.device
        sizes = [len(b) for b in self.blobs]
        n_bytes = sum(sizes)
        # offset of every file's text in the concatenated stream (line offsets refer to that stream)
        self.starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self._starts_list = self.starts.tolist()
        d_text = torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev)
        at = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                  # read-only source buffers: they are only read
            for b in self.blobs:                             # every file goes up on its own: no host-side concatenation
                if b:
                    d_text[at:at + len(b)].copy_(torch.from_numpy(np.frombuffer(b, dtype=np.uint8)))
                at += len(b)
        cap = n_bytes // 48 + 1024                           # LA4Falcon lines are ~75 bytes; grown on demand
        for _ in range(2):
            self.d: Dict[str, "torch.Tensor"] = {k: torch.empty(cap, dtype=torch.int32, device=dev) for k in COLS}
            self.d["flags"] = torch.empty(cap, dtype=torch.uint8, device=dev)
            self.d["off"] = torch.empty(cap, dtype=torch.int64, device=dev)
            self.d["llen"] = torch.empty(cap, dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)
            _lib.check(eng.ctx, lib().fuz_parse_la4falcon(eng.ctx, d_text.data_ptr(), n_bytes, cap, 1 if require_id9 else 0,
                                                           *[self.d[k].data_ptr() for k in COLS], self.d["flags"].data_ptr(),
                                                           self.d["off"].data_ptr(), self.d["llen"].data_ptr()))
            st = eng.status(raise_on_error=False)
            if st.error == _lib.FUZ_E_CAPACITY and st.error_index == 12:
                cap = int(st.reserved[0]) + 16
                continue
            break
        if st.error == _lib.FUZ_E_FORMAT:
            if int(st.reserved[3]) == 3:
                raise FuzError(st.error, "read ids of the overlap lines must be %09d ids (the reference compares them as strings)")
            raise ValueError("malformed LA4Falcon line %d (12+ columns: ids, lengths and coordinates as integers, identity as float)"
                             % st.error_index)
        if st.error:
            eng.status()
        self.n = int(st.reserved[0])
        for k in self.d:
            self.d[k] = self.d[k][:max(self.n, 1)]
        if int(st.reserved[1]):
            self._host_identity()
        ends = torch.from_numpy(np.cumsum(sizes).astype(np.int64)).to(dev)
        self.d["file"] = torch.bucketize(self.d["off"][:self.n], ends, right=True).to(torch.int32) if self.n else torch.zeros(
            1, dtype=torch.int32, device=dev)
        self._a: Optional[Dict[str, np.ndarray]] = None
        del d_text

    @property
    def text(self) -> bytes:
        """The concatenated text (made when the first selected line has to be printed)."""
        if self._text is None:
            self._text = b"".join(self.blobs)
        return self._text

    def _host_identity(self) -> None:
        """float(l[3]) < 90 for the lines the kernel left open (flags bit 7)."""
        import torch
        flags = self.d["flags"][:self.n]
        idx = torch.nonzero(flags >= 128).flatten()
        off, llen = self.d["off"][idx].cpu().numpy(), self.d["llen"][idx].cpu().numpy()
        new = []
        for o, l in zip(off.tolist(), llen.tolist()):
            tok = self.text[o:o + l].split()[3].decode("ascii", "replace")
            if not _PY2_FLOAT.match(tok):
                raise ValueError("could not convert string to float: %r" % tok)
            new.append(0 if float(tok) < 90 else 1)
        cur = flags[idx].cpu().numpy()
        cur = (cur & 0x7E) | np.asarray(new, np.uint8)
        flags[idx] = torch.from_numpy(cur).to(flags.device)

    # ---- host views, made on demand
    @property
    def a(self) -> Dict[str, np.ndarray]:
        if self._a is None:
            self._a = {k: self.d[k][:self.n].cpu().numpy() for k in COLS + ("flags", "off", "llen")}
        return self._a

    @property
    def file(self) -> np.ndarray:
        return self.d["file"][:self.n].cpu().numpy()

    def gather(self, keys: Sequence[str], sel: np.ndarray) -> Dict[str, np.ndarray]:
        """Host copies of some columns for the lines `sel` only."""
        import torch
        idx = torch.from_numpy(np.ascontiguousarray(sel, dtype=np.int64)).to(self.d["q"].device)
        return {k: self.d[k][idx].cpu().numpy() if len(sel) else np.zeros(0, dtype=self.d[k].cpu().numpy().dtype) for k in keys}


    def tokens(self, i: int) -> List[str]:
        o, l = int(self.a["off"][i]), int(self.a["llen"][i])
        k = bisect.bisect_right(self._starts_list, o) - 1
        o -= self._starts_list[k]
        return self.blobs[k][o:o + l].decode("ascii").split()

"""Tiny reader for R's RDX2/XDR serialisation (gzip'd .RData), enough for the reference fixtures.

The reference ships its example inputs as .RData (reference data/Bourne.RData,
data/ATNeu_example.RData, data/SA_cru.RData).  Neither R nor
pyreadr/rdata is available in this image, so this decodes the subset of SEXP types those
files contain.  It is used ONLY by tools/make_golden.py (run in the build container, where
/root/reference exists) to turn the fixtures into small .npz inputs under tests/golden/.

Format reference: R Internals, "Serialization Formats".  Each item starts with a 32-bit
big-endian flags word: bits 0-7 type, bit 8 is-object, bit 9 has-attributes, bit 10 has-tag.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
SPECIALSXP, BUILTINSXP, CHARSXP, LGLSXP, INTSXP, REALSXP = 7, 8, 9, 10, 13, 14
CPLXSXP, STRSXP, DOTSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, RAWSXP, S4SXP = 15, 16, 17, 19, 20, 21, 22, 24, 25
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP = 255, 254, 253, 252, 251
BASENAMESPACE_SXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 250, 249, 248, 247
EMPTYENV_SXP, BASEENV_SXP = 242, 241
ATTRLANGSXP, ATTRLISTSXP = 240, 239


class RObj:
    """A decoded R value: `.value` (numpy array / list / str / None) plus `.attr` dict."""

    def __init__(self, value, attr=None, tag=None):
        self.value = value
        self.attr = attr or {}
        self.tag = tag

    def names(self):
        n = self.attr.get("names")
        return list(n.value) if n is not None else None

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.value[self.names().index(key)]
        return self.value[key]

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


class Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.refs = []

    def i32(self):
        (v,) = struct.unpack_from(">i", self.d, self.p)
        self.p += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def pairlist_to_dict(self, obj):
        out = {}
        while obj is not None and isinstance(obj.value, tuple):
            car, cdr = obj.value
            out[obj.tag] = car
            obj = cdr
        return out

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP, BASENAMESPACE_SXP):
            return RObj(f"<pseudo:{t}>")
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            sym = RObj(name.value if isinstance(name, RObj) else name)
            self.refs.append(sym)
            return sym
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            vals = [self.item() for _ in range(n)]
            o = RObj(("namespace", vals))
            self.refs.append(o)
            return o
        if t == ENVSXP:
            self.i32()  # locked
            o = RObj("<env>")
            self.refs.append(o)
            self.item()  # enclos
            self.item()  # frame
            self.item()  # hashtab
            self.item()  # attrib
            return o
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP, ATTRLANGSXP, ATTRLISTSXP):
            attr = None
            if has_attr or t in (ATTRLANGSXP, ATTRLISTSXP):
                attr = self.item()
            tag = None
            if has_tag:
                tg = self.item()
                tag = tg.value if isinstance(tg, RObj) else tg
            car = self.item()
            cdr = self.item()
            return RObj((car, cdr), tag=tag)
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                return RObj(None)
            s = self.d[self.p:self.p + n].decode("latin-1")
            self.p += n
            return RObj(s)
        if t in (LGLSXP, INTSXP):
            n = self.length()
            v = np.frombuffer(self.d, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            o = RObj(v)
        elif t == REALSXP:
            n = self.length()
            v = np.frombuffer(self.d, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            o = RObj(v)
        elif t == STRSXP:
            n = self.length()
            o = RObj([self.item().value for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            o = RObj([self.item() for _ in range(n)])
        elif t == S4SXP:
            o = RObj("<S4>")
        elif t == RAWSXP:
            n = self.length()
            o = RObj(self.d[self.p:self.p + n])
            self.p += n
        elif t in (SPECIALSXP, BUILTINSXP):
            n = self.i32()
            o = RObj(self.d[self.p:self.p + n].decode())
            self.p += n
        elif t == BCODESXP:
            raise NotImplementedError("byte-code objects are not decoded")
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.p}")
        if has_attr:
            o.attr = self.pairlist_to_dict(self.item())
        return o


def decompress(path: str) -> bytes:
    raw = open(path, "rb").read()
    data = gzip.decompress(raw)
    assert data[:5] == b"RDX2\n", data[:8]
    return data


def load_rdata(path: str) -> dict:
    """Decode a whole .RData file into {object name: RObj}."""
    data = decompress(path)
    r = Reader(data)
    r.p = 5
    assert data[r.p:r.p + 2] == b"X\n"
    r.p += 2
    r.i32(), r.i32(), r.i32()  # format version, writer version, min reader version
    top = r.item()
    return r.pairlist_to_dict(top)


def find_real_vectors(path: str, lengths):
    """Locate REALSXP payloads of the given lengths by scanning for their headers.

    SA_cru.RData holds S4 RasterBrick objects whose closure slots carry byte-code; instead of
    decoding those, find the value arrays by their (type=REALSXP, length) header (SURVEY.md
    Appendix C).  Returns {length: [np.ndarray, ...]} in file order.
    """
    data = decompress(path)
    out = {n: [] for n in lengths}
    for n in lengths:
        pat = struct.pack(">i", n)
        start = 0
        while True:
            i = data.find(pat, start)
            if i < 0:
                break
            start = i + 1
            if i < 4:
                continue
            (flags,) = struct.unpack_from(">i", data, i - 4)
            if (flags & 0xFF) != REALSXP or (flags >> 12) != 0:
                continue
            if i + 4 + 8 * n > len(data):
                continue
            v = np.frombuffer(data, dtype=">f8", count=n, offset=i + 4).astype(np.float64)
            out[n].append((i - 4, v))
    return out

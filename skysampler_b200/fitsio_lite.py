"""Minimal FITS binary-table writer / reader for the sampler's on-disk tables.

The reference drivers persist proposals and scores with ``fitsio.write(fname, df.to_records(), clobber=True)``
(bin/delta_rejection_sampler_v01.py:105,111; bin/epsilon_concentric_sampler_v05.py:162-164,179-181) and
``read_concentric`` loads them back with ``fitsio.read`` (emulator.py:725,732).  ``fitsio`` is not in this
image, so this module writes / reads the same layout — an empty primary HDU followed by ONE ``BINTABLE``
extension with scalar columns (TFORM 'D' float64, 'E' float32, 'K' int64, 'J' int32), big-endian rows
padded to 2880-byte blocks (FITS standard 4.0, §7.3) — so that files written here open in fitsio /
astropy and vice versa.  Only what those two call sites need is implemented.  Byte-for-byte equality
with fitsio's own output is NOT pinned (fitsio cannot be run here); the tests pin the round trip and the
block / card structure.
"""
import os

import numpy as np

BLOCK = 2880
_TFORM = {"f8": "D", "f4": "E", "i8": "K", "i4": "J"}
_DTYPE = {"D": ">f8", "E": ">f4", "K": ">i8", "J": ">i4"}


def _card(key, value=None, comment=""):
    if value is None:
        return ("%-8s" % key).ljust(80)
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
        v = v.ljust(20)
    s = "%-8s= %s" % (key, v)
    if comment:
        s += " / " + comment
    return s[:80].ljust(80)


def _header(cards):
    txt = "".join(cards) + _card("END")
    pad = (-len(txt)) % BLOCK
    return (txt + " " * pad).encode("ascii")


def write(filename, data, clobber=False):
    """Write a numpy record array (e.g. ``DataFrame.to_records()``) as a one-extension FITS file."""
    if os.path.exists(filename) and not clobber:
        raise IOError("file exists: %s (use clobber=True)" % filename)
    data = np.asarray(data)
    names = list(data.dtype.names)
    forms, cols = [], []
    for n in names:
        dt = data.dtype[n]
        kind = "%s%d" % (dt.kind, dt.itemsize)
        if dt.kind == "O" or kind not in _TFORM:
            arr = np.asarray(data[n].tolist())
            kind = "%s%d" % (arr.dtype.kind, arr.dtype.itemsize)
            if kind not in _TFORM:
                raise TypeError("column %r has unsupported dtype %s" % (n, dt))
        else:
            arr = data[n]
        forms.append(_TFORM[kind])
        cols.append(np.ascontiguousarray(arr).astype(_DTYPE[_TFORM[kind]]))
    nrows = len(data)
    rowdt = np.dtype([(n, _DTYPE[f]) for n, f in zip(names, forms)])
    table = np.empty(nrows, dtype=rowdt)
    for n, c in zip(names, cols):
        table[n] = c
    primary = _header([_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", 8),
                       _card("NAXIS", 0), _card("EXTEND", True)])
    cards = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8), _card("NAXIS", 2),
             _card("NAXIS1", rowdt.itemsize, "width of table in bytes"), _card("NAXIS2", nrows, "number of rows"),
             _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(names))]
    for i, (n, f) in enumerate(zip(names, forms), 1):
        cards.append(_card("TTYPE%d" % i, n))
        cards.append(_card("TFORM%d" % i, f))
    body = table.tobytes()
    with open(filename, "wb") as fh:
        fh.write(primary)
        fh.write(_header(cards))
        fh.write(body)
        fh.write(b"\0" * ((-len(body)) % BLOCK))


def _read_header(fh):
    cards = {}
    order = []
    while True:
        block = fh.read(BLOCK)
        if len(block) < BLOCK:
            raise IOError("truncated FITS header")
        for i in range(0, BLOCK, 80):
            c = block[i:i + 80].decode("ascii")
            key = c[:8].strip()
            if key == "END":
                return cards, order
            if c[8:10] == "= ":
                val = c[10:].split(" /")[0].strip()
                if val.startswith("'"):
                    val = val[1:val.rindex("'")].rstrip().replace("''", "'")
                elif val in ("T", "F"):
                    val = (val == "T")
                else:
                    val = int(val) if val.lstrip("+-").isdigit() else float(val)
                cards[key] = val
                order.append(key)


def read(filename, ext=1):
    """Read the first BINTABLE extension into a native-endian numpy record array."""
    with open(filename, "rb") as fh:
        hdr, _ = _read_header(fh)
        if not hdr.get("SIMPLE", False):
            raise IOError("not a FITS file: %s" % filename)
        nax = hdr.get("NAXIS", 0)
        size = abs(hdr.get("BITPIX", 8)) // 8
        for i in range(1, nax + 1):
            size *= hdr["NAXIS%d" % i]
        if nax:
            fh.seek(size + ((-size) % BLOCK), 1)
        hdr, _ = _read_header(fh)
        if hdr.get("XTENSION") != "BINTABLE":
            raise IOError("first extension is not a BINTABLE")
        nf = hdr["TFIELDS"]
        fields = []
        for i in range(1, nf + 1):
            form = str(hdr["TFORM%d" % i]).strip().lstrip("1")
            if form not in _DTYPE:
                raise TypeError("unsupported TFORM %r" % form)
            fields.append((str(hdr["TTYPE%d" % i]).strip(), _DTYPE[form]))
        rowdt = np.dtype(fields)
        if rowdt.itemsize != hdr["NAXIS1"]:
            raise IOError("row width mismatch")
        raw = fh.read(rowdt.itemsize * hdr["NAXIS2"])
        table = np.frombuffer(raw, dtype=rowdt, count=hdr["NAXIS2"])
    native = np.dtype([(n, np.dtype(t).newbyteorder("=")) for n, t in fields])
    out = np.empty(len(table), dtype=native)
    for n, _ in fields:
        out[n] = table[n]
    return out

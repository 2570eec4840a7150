"""Minimal FITS reader / writer for the files GPPupilDemodulation.jl handles.

The reference reads the primary header keywords and the ``METROLOGY`` binary table
through FITSIO/CFITSIO and writes a copy of the whole file with that table and its
header replaced (``FitsUtils.FITScopy!``, src/FitsUtils.jl:95-156;
``main``, src/GPPupilDemodulation.jl:356-424).  No FITS library exists in this
image, so this module implements exactly what that needs, on plain bytes:

* plain, gzipped (``.gz``) and compress'ed (``.Z``) files,
* header parsing (standard and ``HIERARCH`` cards, CONTINUE-less strings),
* BINTABLE layout (``NAXIS1`` = record bytes, ``TFORMn`` widths -> column offsets),
* byte-exact pass-through of every other HDU,
* rewriting one BINTABLE: records replaced, ``TFORMn``/``NAXIS1`` updated, columns
  appended, header cards added.

The table records themselves are never decoded here in whole-file mode: they go to
the GPU as they are (``gppd_submit_fits_rows``) and come back as records.
"""
from __future__ import annotations

import gzip
import re
from dataclasses import dataclass, field

import numpy as np

BLOCK = 2880
CARD = 80

# bytes per element of the TFORM codes that occur in GRAVITY tables
_TFORM_BYTES = {"L": 1, "X": 1, "B": 1, "I": 2, "J": 4, "K": 8, "A": 1, "E": 4, "D": 8,
                "C": 8, "M": 16, "P": 8, "Q": 16}


def _parse_value(text: str):
    s = text.strip()
    if not s:
        return None
    if s[0] == "'":
        end = 1
        out = []
        while end < len(s):          # '' inside a string is an escaped quote
            if s[end] == "'":
                if end + 1 < len(s) and s[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(s[end])
            end += 1
        return "".join(out).rstrip()
    s = s.split("/")[0].strip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace("D", "E"))
        except ValueError:
            return s


def parse_card(card: str):
    """(keyword, value) of one 80-character card; (None, None) for comments/blank."""
    key = card[:8].rstrip()
    if key == "HIERARCH":
        if "=" not in card:
            return None, None
        k, v = card[8:].split("=", 1)
        return k.strip(), _parse_value(v)
    if key in ("", "COMMENT", "HISTORY", "END") or card[8:10] != "= ":
        return None, None
    return key, _parse_value(card[10:])


def format_card(key: str, value, comment: str = "") -> str:
    """One header card; keywords longer than 8 characters (or with spaces) use the ESO
    HIERARCH convention, like CFITSIO does for them."""
    if isinstance(value, bool):
        v = "T" if value else "F"
    elif isinstance(value, (int, np.integer)):
        v = str(int(value))
    elif isinstance(value, (float, np.floating)):
        v = repr(float(value)).upper() if np.isfinite(value) else "'NaN'"
        if "E" not in v and "." not in v and v[0] != "'":
            v += "."
    else:
        v = "'" + str(value).replace("'", "''") + "'"
    if len(key) > 8 or " " in key:
        card = f"HIERARCH {key} = {v}"
    elif v[0] == "'":
        card = f"{key:<8}= {v:<20}"
    else:
        card = f"{key:<8}= {v:>20}"
    if comment:
        card += " / " + comment
    if len(card) > CARD:
        card = card[:CARD]
    return card.ljust(CARD)


@dataclass
class HDU:
    cards: list            # raw 80-character cards (without END)
    data: bytes            # data unit, padded to a multiple of 2880 bytes (None: scan_fits)
    header: dict = field(default_factory=dict)
    hdr_offset: int = -1   # scan_fits: byte offset of the header in the file
    data_offset: int = -1  # scan_fits: byte offset of the data unit
    data_padded: int = 0   # scan_fits: bytes of the data unit including its padding

    @property
    def name(self):
        return self.header.get("EXTNAME", "")

    @property
    def is_bintable(self):
        return self.header.get("XTENSION") == "BINTABLE"


def _data_bytes(h: dict) -> int:
    naxis = int(h.get("NAXIS", 0))
    if naxis == 0:
        return 0
    n = abs(int(h.get("BITPIX", 8))) // 8
    for i in range(1, naxis + 1):
        n *= int(h.get(f"NAXIS{i}", 0))
    n *= int(h.get("GCOUNT", 1))
    return n + int(h.get("PCOUNT", 0))


def unlzw(buf: bytes) -> bytes:
    """Decode a Unix ``compress`` (.Z) stream: LZW with 9..16-bit codes, least significant bit
    first, code 256 = CLEAR in block mode, and the format's quirk that the encoder emits codes
    in groups of eight, so that a change of code width (or a CLEAR) skips to the next multiple
    of ``8 * width`` bits counted from where that width began.  CFITSIO reads such files
    transparently; the reference lists ``fits.Z`` among its suffixes
    (src/GPPupilDemodulation.jl:14)."""
    if len(buf) < 3 or buf[0] != 0x1F or buf[1] != 0x9D:
        raise ValueError("not a compress (.Z) stream")
    maxbits, block = buf[2] & 0x1F, bool(buf[2] & 0x80)
    if not 9 <= maxbits <= 16 or buf[2] & 0x60:
        raise ValueError("unsupported compress (.Z) flags")
    first_free = 257 if block else 256
    table = [bytes([i]) for i in range(256)] + ([b""] if block else [])
    data = buf[3:] + b"\0\0\0\0"
    total = 8 * (len(buf) - 3)
    out = bytearray()
    bits, mark, pos, prev = 9, 0, 0, None
    while True:
        full = (1 << maxbits) if bits == maxbits else (1 << bits) - 1
        if len(table) > full and bits < maxbits:       # the next code is one bit wider
            group = 8 * bits
            pos += -(pos - mark) % group
            bits += 1
            mark = pos
        if pos + bits > total:
            break
        code = (int.from_bytes(data[pos >> 3:(pos >> 3) + 4], "little") >> (pos & 7)) & ((1 << bits) - 1)
        pos += bits
        if block and code == 256:                       # CLEAR: back to 9-bit codes
            group = 8 * bits
            pos += -(pos - mark) % group
            del table[first_free:]
            bits, mark, prev = 9, pos, None
            continue
        if prev is None:
            if code > 255:
                raise ValueError("corrupt compress (.Z) stream")
            entry = table[code]
        elif code < len(table):
            entry = table[code]
        elif code == len(table):
            entry = prev + prev[:1]
        else:
            raise ValueError("corrupt compress (.Z) stream")
        out += entry
        if prev is not None and len(table) < (1 << maxbits):
            table.append(prev + entry[:1])
        prev = entry
    return bytes(out)


def read_fits(path: str) -> list:
    """All HDUs of a FITS file (plain, gzipped ``.gz`` or compress'ed ``.Z``) as raw bytes +
    parsed headers."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        buf = fh.read()
    if str(path).endswith(".Z"):
        buf = unlzw(buf)
    view = memoryview(buf)      # data units are views: no copy of a 30 MB table per file
    hdus, pos = [], 0
    while pos + BLOCK <= len(buf):
        cards, header, done = [], {}, False
        while not done:
            block = buf[pos:pos + BLOCK].decode("ascii", "replace")
            pos += BLOCK
            for i in range(0, BLOCK, CARD):
                card = block[i:i + CARD]
                if card.startswith("END") and card[3:].strip() == "":
                    done = True
                    break
                cards.append(card)
                k, v = parse_card(card)
                if k is not None and k not in header:
                    header[k] = v
            if pos > len(buf):
                raise ValueError("truncated FITS header")
        nbytes = _data_bytes(header)
        padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
        hdus.append(HDU(cards, view[pos:pos + padded], header))
        pos += padded
    return hdus


def scan_fits(path: str) -> list:
    """The HDUs of a plain (uncompressed) FITS file WITHOUT their data: headers are read,
    data units are skipped (``hdr_offset`` / ``data_offset`` / ``data_padded`` say where they
    are).  What the native record path needs (``gppd_file_submit`` reads the records itself)."""
    hdus = []
    with open(path, "rb") as fh:
        fh.seek(0, 2)
        size = fh.tell()
        pos = 0
        while pos + BLOCK <= size:
            start = pos
            cards, header, done = [], {}, False
            while not done:
                fh.seek(pos)
                block = fh.read(BLOCK).decode("ascii", "replace")
                if len(block) < BLOCK:
                    raise ValueError("truncated FITS header")
                pos += BLOCK
                for i in range(0, BLOCK, CARD):
                    card = block[i:i + CARD]
                    if card.startswith("END") and card[3:].strip() == "":
                        done = True
                        break
                    cards.append(card)
                    k, v = parse_card(card)
                    if k is not None and k not in header:
                        header[k] = v
            nbytes = _data_bytes(header)
            padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
            hdus.append(HDU(cards, None, header, hdr_offset=start, data_offset=pos, data_padded=padded))
            pos += padded
    return hdus


def _header_bytes(cards) -> bytes:
    text = "".join(c.ljust(CARD)[:CARD] for c in cards) + "END".ljust(CARD)
    text += " " * (-len(text) % BLOCK)
    return text.encode("ascii")


def write_fits(path: str, hdus) -> None:
    with open(path, "wb") as fh:
        for h in hdus:
            fh.write(_header_bytes(h.cards))
            fh.write(h.data)
            fh.write(b"\0" * (-len(h.data) % BLOCK))


def tform_bytes(tform: str) -> int:
    m = re.match(r"\s*(\d*)([A-Z])", str(tform))
    if not m:
        raise ValueError(f"unsupported TFORM {tform!r}")
    return (int(m.group(1)) if m.group(1) else 1) * _TFORM_BYTES[m.group(2)]


def bintable_layout(header: dict):
    """(record bytes, rows, {column name: (byte offset, TFORM, index)})."""
    cols, off = {}, 0
    for i in range(1, int(header["TFIELDS"]) + 1):
        tform = header[f"TFORM{i}"]
        cols[str(header.get(f"TTYPE{i}", f"COL{i}")).strip()] = (off, str(tform).strip(), i)
        off += tform_bytes(tform)
    if off != int(header["NAXIS1"]):
        raise ValueError("TFORM widths do not add up to NAXIS1")
    return int(header["NAXIS1"]), int(header["NAXIS2"]), cols


def _set_card(cards, key, value, comment=""):
    new = format_card(key, value, comment)
    for i, c in enumerate(cards):
        if parse_card(c)[0] == key:
            cards[i] = new
            return
    cards.append(new)


def replace_bintable_cards(hdu: HDU, row_bytes: int, nrows: int, tform_changes=None, new_columns=(),
                           new_keys=()) -> list:
    """The header cards of BINTABLE ``hdu`` after its records have been replaced by ``nrows``
    records of ``row_bytes`` bytes (see replace_bintable)."""
    cards = list(hdu.cards)
    header = dict(hdu.header)
    _, _, cols = bintable_layout(header)
    index = {}                      # keyword -> position of its (first) card: one parse per card
    for i, c in enumerate(cards):
        k = parse_card(c)[0]
        if k is not None and k not in index:
            index[k] = i

    def put(key, value):
        new = format_card(key, value)
        if key in index:
            cards[index[key]] = new
        else:
            index[key] = len(cards)
            cards.append(new)

    for name, tf in (tform_changes or {}).items():
        put(f"TFORM{cols[name][2]}", tf)
    nf = int(header["TFIELDS"])
    for name, tf, unit in new_columns:
        nf += 1
        put(f"TTYPE{nf}", name)
        put(f"TFORM{nf}", tf)
        if unit:
            put(f"TUNIT{nf}", unit)
    put("TFIELDS", nf)
    put("NAXIS1", int(row_bytes))
    put("NAXIS2", int(nrows))
    for k, v in new_keys:
        put(k, v)
    # keep the mandatory keywords in their mandatory order: only values were edited
    # or cards appended, so the order of the original header is preserved
    return cards


def header_bytes(cards) -> bytes:
    """The header unit (cards + END, padded to 2880 bytes) as it is stored in a file."""
    return _header_bytes(cards)


def replace_bintable(hdu: HDU, records: np.ndarray, tform_changes=None, new_columns=(),
                     new_keys=()) -> HDU:
    """A copy of BINTABLE ``hdu`` whose records are ``records`` (uint8, rows x bytes).

    tform_changes: {column name: new TFORM} for columns whose width changed in place;
    new_columns: [(name, TFORM, unit or None)] appended after the existing ones (their
    bytes are already at the end of ``records``); new_keys: [(keyword, value)] header
    cards to add or replace (reference: the DEMODULATION keys and PROCSOFT,
    src/GPPupilDemodulation.jl:174-189,252)."""
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    cards = replace_bintable_cards(hdu, int(rec.shape[1]), int(rec.shape[0]), tform_changes, new_columns,
                                   new_keys)
    out = HDU(cards, rec.reshape(-1).data, {})      # a view of the records, not another copy
    for c in cards:
        k, v = parse_card(c)
        if k is not None and k not in out.header:
            out.header[k] = v
    return out


def make_bintable(name: str, columns, keys=()) -> HDU:
    """A new BINTABLE HDU from big-endian column arrays [(name, TFORM, unit, ndarray)]
    (used by the synthetic-file generator of the tests and the bench)."""
    n = columns[0][3].shape[0]
    recs = np.concatenate([np.ascontiguousarray(a).reshape(n, -1).view(np.uint8) for *_, a in columns],
                          axis=1)
    cards = [format_card("XTENSION", "BINTABLE"), format_card("BITPIX", 8), format_card("NAXIS", 2),
             format_card("NAXIS1", int(recs.shape[1])), format_card("NAXIS2", int(n)),
             format_card("PCOUNT", 0), format_card("GCOUNT", 1),
             format_card("TFIELDS", len(columns))]
    for i, (cname, tform, unit, _) in enumerate(columns, 1):
        cards += [format_card(f"TTYPE{i}", cname), format_card(f"TFORM{i}", tform)]
        if unit:
            cards.append(format_card(f"TUNIT{i}", unit))
    cards.append(format_card("EXTNAME", name))
    cards += [format_card(k, v) for k, v in keys]
    hdu = HDU(cards, recs.tobytes(), {})
    for c in cards:
        k, v = parse_card(c)
        if k is not None:
            hdu.header.setdefault(k, v)
    return hdu


def make_primary(keys) -> HDU:
    cards = [format_card("SIMPLE", True), format_card("BITPIX", 8), format_card("NAXIS", 0),
             format_card("EXTEND", True)] + [format_card(k, v) for k, v in keys]
    hdu = HDU(cards, b"", {})
    for c in cards:
        k, v = parse_card(c)
        if k is not None:
            hdu.header.setdefault(k, v)
    return hdu

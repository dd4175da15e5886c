"""Minimal OpenEXR scanline reader for tests (NONE / ZIPS / ZIP compression, HALF / FLOAT channels), written
from the public file-layout specification. Used to read back what adypt_write_exr produces and, when the
reference tree is present, the reference's own showcase renders (written by tinyexr)."""
import struct
import zlib

import numpy as np


def _cstr(buf, pos):
    end = buf.index(b"\0", pos)
    return buf[pos:end].decode(), end + 1


def read_exr(path):
    buf = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", buf, 0)
    assert magic == 20000630, "not an OpenEXR file"
    assert (version & 0xFF) == 2 and not (version & 0x200), "single-part scanline files only"
    pos = 8
    attrs = {}
    while buf[pos] != 0:
        name, pos = _cstr(buf, pos)
        typ, pos = _cstr(buf, pos)
        (size,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        attrs[name] = (typ, buf[pos:pos + size])
        pos += size
    pos += 1
    channels = []
    cb = attrs["channels"][1]
    p = 0
    while cb[p] != 0:
        nm, p = _cstr(cb, p)
        ptype, = struct.unpack_from("<I", cb, p)
        p += 16
        channels.append((nm, ptype))
    comp = attrs["compression"][1][0]
    xmin, ymin, xmax, ymax = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = xmax - xmin + 1, ymax - ymin + 1
    lines_per_chunk = {0: 1, 2: 1, 3: 16}[comp]
    n_chunks = (h + lines_per_chunk - 1) // lines_per_chunk
    offsets = struct.unpack_from("<%dQ" % n_chunks, buf, pos)
    bpc = {1: 2, 2: 4}
    line_bytes = sum(bpc[t] for _, t in channels) * w
    out = {nm: np.zeros((h, w), dtype=np.float32) for nm, _ in channels}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        data = buf[off + 8: off + 8 + size]
        nlines = min(lines_per_chunk, ymax + 1 - y)
        raw_size = line_bytes * nlines
        if comp != 0 and size < raw_size:
            d = np.frombuffer(zlib.decompress(data), dtype=np.uint8).astype(np.int64)
            d = (np.cumsum(d - np.concatenate([[0], np.full(d.size - 1, 128)])) & 0xFF).astype(np.uint8)  # undo predictor
            half = (raw_size + 1) // 2
            raw = np.empty(raw_size, dtype=np.uint8)
            raw[0::2] = d[:half]
            raw[1::2] = d[half:]
            data = raw.tobytes()
        p = 0
        for ln in range(nlines):
            for nm, t in channels:
                nb = bpc[t] * w
                arr = np.frombuffer(data, dtype=np.float16 if t == 1 else np.float32, count=w, offset=p)
                out[nm][y - ymin + ln] = arr.astype(np.float32)
                p += nb
    return dict(width=w, height=h, channels=channels, compression=comp, data=out, attrs=attrs)

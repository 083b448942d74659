"""A small FLAC ENCODER for the tests (test infrastructure, not product code): there is no FLAC tool in this image, so
the ingest tests make their own streams.  It follows the published format (RFC 9639): STREAMINFO, frames with a
UTF-8 coded frame number, CRC-8 / CRC-16, CONSTANT / VERBATIM / FIXED / LPC subframes, partitioned Rice residuals
(4- and 5-bit parameters, escape partitions), wasted bits, and the four stereo modes.  FLAC is lossless: whatever a
correct decoder makes of these streams is the PCM that went in, which is what the tests compare against.

`encode(pcm [frames, channels] int32, rate, bps, ...)` returns the file bytes; `style` picks which features a frame uses so
that every decoder path is exercised, not which compresses best."""
from __future__ import annotations

import numpy as np


class BitWriter:
    def __init__(self):
        self.buf = bytearray(); self.acc = 0; self.n = 0

    def put(self, value: int, bits: int):
        if bits == 0:
            return
        value &= (1 << bits) - 1
        self.acc = (self.acc << bits) | value; self.n += bits
        while self.n >= 8:
            self.n -= 8
            self.buf.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def put_signed(self, value: int, bits: int):
        self.put(value & ((1 << bits) - 1), bits)

    def unary(self, q: int):                       # q zeros then a one
        while q >= 32:
            self.put(0, 32); q -= 32
        self.put(1, q + 1)

    def align(self):
        if self.n:
            self.put(0, 8 - self.n)

    def bytes(self) -> bytes:
        assert self.n == 0
        return bytes(self.buf)


def crc8(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


_CRC16 = []
for _i in range(256):
    _c = _i << 8
    for _ in range(8):
        _c = ((_c << 1) ^ 0x8005) & 0xFFFF if _c & 0x8000 else (_c << 1) & 0xFFFF
    _CRC16.append(_c)


def crc16(data: bytes) -> int:
    c = 0
    for b in data:
        c = ((c << 8) & 0xFFFF) ^ _CRC16[(c >> 8) ^ b]
    return c


def utf8_number(v: int) -> bytes:
    if v < 0x80:
        return bytes([v])
    n = 2
    while v >= (1 << (5 * n + 1)) and n < 7:          # n bytes carry 5n+1 bits (n >= 2)
        n += 1
    out = []
    for _ in range(n - 1):
        out.append(0x80 | (v & 0x3F)); v >>= 6
    lead = (0xFF << (8 - n)) & 0xFF
    out.append(lead | v)
    return bytes(reversed(out))


def zigzag(r: np.ndarray) -> np.ndarray:
    r = r.astype(np.int64)
    return np.where(r >= 0, r << 1, ((-r) << 1) - 1)


def best_rice(u: np.ndarray, max_k: int) -> int:
    best, bk = None, 0
    for k in range(0, max_k):
        cost = int((u >> k).sum()) + (k + 1) * u.size
        if best is None or cost < best:
            best, bk = cost, k
    return bk


def write_residual(w: BitWriter, res: np.ndarray, order: int, blocksize: int, part_order: int, five_bit: bool, escape_first: bool):
    w.put(1 if five_bit else 0, 2)
    w.put(part_order, 4)
    pbits, esc = (5, 31) if five_bit else (4, 15)
    pos = 0
    for p in range(1 << part_order):
        n = (blocksize >> part_order) - (order if p == 0 else 0)
        part = res[pos: pos + n]; pos += n
        if escape_first and p == 0 and n > 0:
            w.put(esc, pbits)
            nb = max(1, int(np.abs(part.astype(np.int64)).max()).bit_length() + 1)
            w.put(nb, 5)
            for v in part:
                w.put_signed(int(v), nb)
            continue
        u = zigzag(part)
        k = best_rice(u, esc) if n else 0
        w.put(k, pbits)
        for v in u:
            v = int(v)
            w.unary(v >> k)
            w.put(v & ((1 << k) - 1), k)
    assert pos == res.size


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def predict_residual(x: np.ndarray, coefs, shift: int) -> np.ndarray:
    order = len(coefs)
    x = x.astype(np.int64)
    pred = np.zeros(x.size - order, np.int64)
    for j, c in enumerate(coefs):
        pred += int(c) * x[order - 1 - j: x.size - 1 - j]
    return x[order:] - (pred >> shift)


def lpc_coefs(x: np.ndarray, order: int, precision: int):
    """Quantised LPC by autocorrelation + Levinson (any stable-ish predictor will do: the residual absorbs the rest)."""
    xf = x.astype(np.float64) * np.hanning(x.size)
    r = np.array([np.dot(xf[: xf.size - k], xf[k:]) for k in range(order + 1)])
    if r[0] <= 0:
        return [0] * order, 0
    a = np.zeros(order + 1); a[0] = 1.0; e = r[0]
    for i in range(1, order + 1):
        acc = r[i] + np.dot(a[1:i], r[i - 1:0:-1])
        k = -acc / e if e > 0 else 0.0
        a_new = a.copy()
        a_new[1:i] = a[1:i] + k * a[i - 1:0:-1]
        a_new[i] = k
        a = a_new; e *= (1 - k * k)
    c = -a[1:]
    cmax = np.abs(c).max()
    if cmax == 0:
        return [0] * order, 0
    shift = max(0, min(15, precision - 1 - int(np.floor(np.log2(cmax))) - 1))
    q = np.clip(np.round(c * (1 << shift)), -(1 << (precision - 1)), (1 << (precision - 1)) - 1).astype(np.int64)
    return [int(v) for v in q], shift


def write_subframe(w: BitWriter, x: np.ndarray, bps: int, kind: str, part_order: int, five_bit: bool, escape_first: bool,
                   lpc_order: int = 8, lpc_precision: int = 12):
    blocksize = x.size
    x = x.astype(np.int64)
    wasted = 0
    if kind != "constant" and np.any(x != 0):
        allor = int(np.bitwise_or.reduce(x))
        while wasted < bps - 1 and not (allor >> wasted) & 1:
            wasted += 1
    if wasted:
        x = x >> wasted
    sb = bps - wasted
    w.put(0, 1)
    if kind == "constant":
        assert np.all(x == x[0])
        w.put(0, 6); w.put(0, 1); w.put_signed(int(x[0]), bps)
        return
    def wasted_flag():
        if wasted:
            w.put(1, 1); w.unary(wasted - 1)
        else:
            w.put(0, 1)
    if kind == "verbatim":
        w.put(1, 6); wasted_flag()
        for v in x:
            w.put_signed(int(v), sb)
        return
    if kind.startswith("fixed"):
        order = int(kind[5:])
        w.put(0b001000 | order, 6); wasted_flag()
        for v in x[:order]:
            w.put_signed(int(v), sb)
        res = predict_residual(x, FIXED[order], 0)
    else:
        order = lpc_order
        coefs, shift = lpc_coefs(x, order, lpc_precision)
        w.put(0b100000 | (order - 1), 6); wasted_flag()
        for v in x[:order]:
            w.put_signed(int(v), sb)
        w.put(lpc_precision - 1, 4); w.put_signed(shift, 5)
        for c in coefs:
            w.put_signed(c, lpc_precision)
        res = predict_residual(x, coefs, shift)
    assert np.abs(res).max(initial=0) < (1 << 31)
    while part_order > 0 and ((blocksize >> part_order) << part_order != blocksize or (blocksize >> part_order) <= order):
        part_order -= 1
    write_residual(w, res, order, blocksize, part_order, five_bit, escape_first)


BLOCK_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}
RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
BPS_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def encode_frame(block: np.ndarray, frame_no: int, rate: int, bps: int, style: dict, first_sample: int | None = None) -> bytes:
    """block: [n, channels] int64.  first_sample: variable-blocksize stream (the header carries the sample number)."""
    n, ch = block.shape
    w = BitWriter()
    w.put(0b11111111111110, 14); w.put(0, 1); w.put(1 if first_sample is not None else 0, 1)
    bcode = BLOCK_CODES.get(n)
    if bcode is None:
        bcode = 6 if n <= 256 else 7
    w.put(bcode, 4)
    rcode = RATE_CODES.get(rate, 0) if style.get("rate_in_header", True) else 0
    if style.get("rate_explicit"):
        rcode = 12 if rate % 1000 == 0 and rate // 1000 < 256 else (13 if rate < 65536 else 14)
    w.put(rcode, 4)
    mode = style.get("stereo", "independent") if ch == 2 else "independent"
    w.put({"independent": ch - 1, "left_side": 8, "right_side": 9, "mid_side": 10}[mode], 4)
    w.put(BPS_CODES[bps] if style.get("bps_in_header", True) else 0, 3); w.put(0, 1)
    for b in utf8_number(frame_no if first_sample is None else first_sample):
        w.put(b, 8)
    if bcode == 6:
        w.put(n - 1, 8)
    elif bcode == 7:
        w.put(n - 1, 16)
    if rcode == 12:
        w.put(rate // 1000, 8)
    elif rcode == 13:
        w.put(rate, 16)
    elif rcode == 14:
        w.put(rate // 10, 16)
    w.put(crc8(bytes(w.buf)), 8)
    chans = [block[:, c] for c in range(ch)]
    bpss = [bps] * ch
    if mode == "left_side":
        chans = [block[:, 0], block[:, 0] - block[:, 1]]; bpss = [bps, bps + 1]
    elif mode == "right_side":
        chans = [block[:, 0] - block[:, 1], block[:, 1]]; bpss = [bps + 1, bps]
    elif mode == "mid_side":
        chans = [(block[:, 0] + block[:, 1]) >> 1, block[:, 0] - block[:, 1]]; bpss = [bps, bps + 1]
    kinds = style.get("kinds", ["fixed2"])
    for c, x in enumerate(chans):
        kind = kinds[(frame_no + c) % len(kinds)]
        if kind == "constant" and not np.all(x == x[0]):
            kind = "fixed1"
        write_subframe(w, x, bpss[c], kind, style.get("part_order", 3), style.get("five_bit", False),
                       style.get("escape_first", False) and frame_no % 3 == 0, style.get("lpc_order", 8), style.get("lpc_precision", 12))
    w.align()
    body = w.bytes()
    return body + crc16(body).to_bytes(2, "big")


def encode(pcm: np.ndarray, rate: int, bps: int, blocksize: int = 4096, style: dict | None = None, junk_metadata: bool = True,
           return_frames: bool = False):
    """pcm: [frames, channels] integers of `bps` bits.  Fixed-blocksize stream (the last block may be short) unless
    style["variable_blocks"] is set.  return_frames: also the list of (byte offset, bytes) of every frame."""
    style = dict(style or {})
    pcm = np.asarray(pcm).astype(np.int64)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    total, ch = pcm.shape
    sizes = style.get("variable_blocks")
    if sizes:                                        # variable-blocksize stream: block sizes cycle through `sizes`
        frames, pos, i = [], 0, 0
        while pos < total:
            n = min(sizes[i % len(sizes)], total - pos)
            frames.append(encode_frame(pcm[pos: pos + n], i, rate, bps, style, first_sample=pos))
            pos += n; i += 1
        min_bs, max_bs = min(sizes), max(sizes)
    else:
        frames = [encode_frame(pcm[s: s + blocksize], i, rate, bps, style) for i, s in enumerate(range(0, total, blocksize))]
        min_bs = max_bs = blocksize
    info = BitWriter()
    info.put(min_bs, 16); info.put(max_bs, 16)
    info.put(min(len(f) for f in frames) if frames else 0, 24); info.put(max(len(f) for f in frames) if frames else 0, 24)
    info.put(rate, 20); info.put(ch - 1, 3); info.put(bps - 1, 5); info.put(total, 36)
    for _ in range(16):
        info.put(0, 8)                                  # MD5 not computed (all zero = unknown)
    out = bytearray(b"fLaC")
    si = info.bytes()
    out += bytes([0x00 if junk_metadata else 0x80]) + len(si).to_bytes(3, "big") + si
    if junk_metadata:                                    # a PADDING block (type 1) and a VORBIS_COMMENT-like block (type 4), the latter last
        out += bytes([0x01]) + (10).to_bytes(3, "big") + bytes(10)
        vc = b"\x04\x00\x00\x00test" + b"\x00\x00\x00\x00"
        out += bytes([0x84]) + len(vc).to_bytes(3, "big") + vc
    where = []
    for f in frames:
        where.append((len(out), len(f)))
        out += f
    return (bytes(out), where) if return_frames else bytes(out)

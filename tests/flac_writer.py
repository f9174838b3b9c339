"""Test-side FLAC ENCODER (RFC 9639), written only to produce inputs for the native decoder's tests: no FLAC
encoder or decoder exists in the offline image.  Covers every bitstream feature the decoder implements: CONSTANT /
VERBATIM / FIXED (orders 0-4) / LPC subframes, Rice and Rice2 residuals with several partition orders and escape
partitions, wasted bits, independent / left-side / right-side / mid-side stereo, fixed and custom block sizes,
CRC-8 / CRC-16, and the STREAMINFO MD5 of the PCM (which the tests re-check on the decoder's output)."""
import hashlib
import struct

import numpy as np


class BitWriter:
    def __init__(self):
        self.acc, self.nbits, self.out = 0, 0, bytearray()

    def write(self, value, bits):
        if bits == 0:
            return
        value &= (1 << bits) - 1
        self.acc = (self.acc << bits) | value
        self.nbits += bits
        while self.nbits >= 8:
            self.nbits -= 8
            self.out.append((self.acc >> self.nbits) & 0xFF)
        self.acc &= (1 << self.nbits) - 1

    def unary(self, q):
        while q >= 32:
            self.write(0, 32)
            q -= 32
        self.write(1, q + 1)

    def align(self):
        if self.nbits:
            self.write(0, 8 - self.nbits)

    def bytes(self):
        assert self.nbits == 0
        return bytes(self.out)


def crc8(data):
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(data):
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def _utf8_number(n):
    if n < 0x80:
        return bytes([n])
    out, prefix_len = [], 2
    while True:
        cap = 1 << (6 * (prefix_len - 1) + (7 - prefix_len))
        if n < cap:
            break
        prefix_len += 1
    for i in range(prefix_len - 1):
        out.append(0x80 | (n & 0x3F))
        n >>= 6
    first = ((0xFF << (8 - prefix_len)) & 0xFF) | n
    return bytes([first] + out[::-1])


def _rice_bits(res, k):
    u = [(r << 1) if r >= 0 else ((-r << 1) - 1) for r in res]
    return sum((x >> k) + 1 + k for x in u)


def _write_residual(bw, res, blocksize, order, porder, method, force_escape=False):
    bw.write(method, 2)
    bw.write(porder, 4)
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    idx = 0
    for p in range(1 << porder):
        count = (blocksize - order) if porder == 0 else ((blocksize >> porder) - (order if p == 0 else 0))
        part = res[idx: idx + count]
        idx += count
        if force_escape and p % 2 == 1:
            raw = max([int(abs(int(r))).bit_length() + 1 for r in part] + [1])
            bw.write(esc, pbits)
            bw.write(raw, 5)
            for r in part:
                bw.write(int(r), raw)
            continue
        kmax = esc - 1
        k = min(range(kmax + 1), key=lambda kk: _rice_bits([int(r) for r in part], kk)) if len(part) else 0
        bw.write(k, pbits)
        for r in part:
            r = int(r)
            u = (r << 1) if r >= 0 else ((-r << 1) - 1)
            bw.unary(u >> k)
            bw.write(u & ((1 << k) - 1), k)
    assert idx == len(res)


FIXED_COEF = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _write_subframe(bw, x, bps, kind, opts):
    """x: python ints of one channel of one block."""
    n = len(x)
    wasted = opts.get("wasted", 0)
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
        bps -= wasted
    bw.write(0, 1)
    if kind == "constant":
        assert len(set(x)) == 1
        bw.write(0, 6)
    elif kind == "verbatim":
        bw.write(1, 6)
    elif kind == "fixed":
        bw.write(8 + opts["order"], 6)
    else:
        bw.write(32 + opts["order"] - 1, 6)
    if wasted:
        bw.write(1, 1)
        bw.unary(wasted - 1)
    else:
        bw.write(0, 1)
    if kind == "constant":
        bw.write(x[0], bps)
    elif kind == "verbatim":
        for v in x:
            bw.write(v, bps)
    else:
        order = opts["order"]
        if kind == "fixed":
            coef, shift = FIXED_COEF[order], 0
        else:
            coef, shift, prec = opts["coef"], opts["shift"], opts["precision"]
        for v in x[:order]:
            bw.write(v, bps)
        if kind == "lpc":
            bw.write(prec - 1, 4)
            bw.write(shift, 5)
            for c in coef:
                bw.write(c, prec)
        res = []
        for i in range(order, n):
            pred = sum(coef[j] * x[i - 1 - j] for j in range(order)) >> shift
            res.append(x[i] - pred)
        _write_residual(bw, res, n, order, opts.get("porder", 0), opts.get("method", 0), opts.get("escape", False))


def quantize_lpc(x, order, precision=12):
    """Least-squares predictor of the block, quantised the way an encoder does (coef * 2^shift, rounded)."""
    x = np.asarray(x, dtype=np.float64)
    A = np.stack([x[order - 1 - j: len(x) - 1 - j] for j in range(order)], axis=1)
    sol = np.linalg.lstsq(A, x[order:], rcond=None)[0]
    cmax = max(np.abs(sol).max(), 1e-9)
    shift = min(15, max(0, precision - 1 - int(np.ceil(np.log2(cmax + 1e-12))) - 1))
    q = np.clip(np.round(sol * (1 << shift)), -(1 << (precision - 1)), (1 << (precision - 1)) - 1).astype(int)
    return [int(v) for v in q], shift, precision


def encode_flac(pcm, sample_rate, bps, blocksize=4096, plan=None, id3=False):
    """pcm: int array [channels, n].  plan(block_index, channel) -> (kind, opts) and plan.stereo(block_index) ->
    0 independent / 8 left-side / 9 right-side / 10 mid-side.  Returns the file's bytes."""
    pcm = np.asarray(pcm, dtype=np.int64)
    C, n = pcm.shape
    plan = plan or (lambda b, c: ("fixed", {"order": 2, "porder": 2}))
    stereo = getattr(plan, "stereo", lambda b: 0)
    frames = bytearray()
    nblocks = -(-n // blocksize)
    min_fs, max_fs = 1 << 30, 0
    for b in range(nblocks):
        blk = pcm[:, b * blocksize: (b + 1) * blocksize]
        bs = blk.shape[1]
        hdr = BitWriter()
        hdr.write(0x3FFE, 14); hdr.write(0, 1); hdr.write(0, 1)         # sync, reserved, fixed blocksize stream
        std = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}
        bs_code = std.get(bs, 6 if bs <= 256 else 7)
        hdr.write(bs_code, 4)
        sr_codes = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
        sr_code = sr_codes.get(sample_rate, 13 if sample_rate < 65536 else 0)
        hdr.write(sr_code, 4)
        mode = stereo(b) if C == 2 else 0
        hdr.write(mode if mode >= 8 else C - 1, 4)
        hdr.write({8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}.get(bps, 0), 3)
        hdr.write(0, 1)
        for byte in _utf8_number(b):
            hdr.write(byte, 8)
        if bs_code == 6:
            hdr.write(bs - 1, 8)
        elif bs_code == 7:
            hdr.write(bs - 1, 16)
        if sr_code == 13:
            hdr.write(sample_rate, 16)
        head = hdr.bytes()
        bw = BitWriter()
        for byte in head:
            bw.write(byte, 8)
        bw.write(crc8(head), 8)
        chans = [[int(v) for v in blk[c]] for c in range(C)]
        bits = [bps] * C
        if mode == 8:
            chans[1] = [l - r for l, r in zip(chans[0], chans[1])]; bits[1] += 1
        elif mode == 9:
            chans[0] = [l - r for l, r in zip(chans[0], chans[1])]; bits[0] += 1
        elif mode == 10:
            mid = [(l + r) >> 1 for l, r in zip(chans[0], chans[1])]
            side = [l - r for l, r in zip(chans[0], chans[1])]
            chans, bits = [mid, side], [bps, bps + 1]
        for c in range(C):
            kind, opts = plan(b, c)
            opts = dict(opts)
            if kind == "lpc" and "coef" not in opts:
                w = opts.get("wasted", 0)
                opts["coef"], opts["shift"], opts["precision"] = quantize_lpc([v >> w for v in chans[c]], opts["order"], opts.get("precision", 12))
            if kind in ("fixed", "lpc") and opts.get("porder", 0) > 0:
                while opts["porder"] > 0 and (bs % (1 << opts["porder"]) != 0 or (bs >> opts["porder"]) <= opts["order"]):
                    opts["porder"] -= 1
            _write_subframe(bw, chans[c], bits[c], kind, opts)
        bw.align()
        body = bw.bytes()
        frame = body + struct.pack(">H", crc16(body))
        min_fs, max_fs = min(min_fs, len(frame)), max(max_fs, len(frame))
        frames += frame
    bytes_per = (bps + 7) // 8
    md5 = hashlib.md5(b"".join(int(v).to_bytes(bytes_per, "little", signed=True) for v in pcm.T.reshape(-1))).digest()
    si = BitWriter()
    si.write(blocksize, 16); si.write(blocksize, 16); si.write(min_fs, 24); si.write(max_fs, 24)
    si.write(sample_rate, 20); si.write(C - 1, 3); si.write(bps - 1, 5); si.write(n, 36)
    out = bytearray()
    if id3:
        out += b"ID3\x04\x00\x00" + bytes([0, 0, 0, 10]) + b"\x00" * 10
    out += b"fLaC"
    out += bytes([0x00, 0, 0, 34]) + si.bytes() + md5
    out += bytes([0x84, 0, 0, 4]) + b"test"                                  # a last VORBIS_COMMENT-typed block, ignored
    return bytes(out + frames)

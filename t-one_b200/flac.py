"""Minimal pure-Python FLAC reader: enough of the format to decode the reference's example recordings
(reference: tone/demo/audio_examples/*.flac - 8 kHz, 16-bit, mono, fixed block size 576) without miniaudio / torchcodec,
which this image does not have.  The reference reads them with miniaudio (tone/demo/read_audio.py:25-53) and feeds
``StreamingCTCPipeline.forward_offline`` an int32 array (tone/pipeline.py:174-203); ``read_flac`` returns the same.

Supports: STREAMINFO, fixed and variable block sizes, 1-8 channels with independent / left-side / right-side / mid-side
decorrelation, 4-32 bits per sample, CONSTANT / VERBATIM / FIXED / LPC subframes, wasted bits, Rice partitions (4- and
5-bit parameters, escape codes).  Frame CRCs are not checked; instead ``verify=True`` checks the decoded PCM against the
MD5 signature in STREAMINFO, which is the stronger end-to-end check.
"""
from __future__ import annotations

import hashlib
import struct
from typing import List, Tuple

import numpy as np

_LEADING_ZEROS = bytes(8 - i.bit_length() for i in range(256))      # leading zero bits of a byte value


class _Bits:
    __slots__ = ("d", "pos")

    def __init__(self, data: bytes, byte_pos: int):
        self.d, self.pos = data, byte_pos * 8

    def read(self, n: int) -> int:
        if n == 0:
            return 0
        p = self.pos
        b0, b1 = p >> 3, (p + n + 7) >> 3
        v = int.from_bytes(self.d[b0:b1], "big")
        self.pos = p + n
        return (v >> (b1 * 8 - p - n)) & ((1 << n) - 1)

    def read_signed(self, n: int) -> int:
        v = self.read(n)
        return v - (1 << n) if n and v >> (n - 1) else v

    def unary(self) -> int:
        """Number of 0 bits before the next 1 bit (which is consumed)."""
        d, p = self.d, self.pos
        byte, off = p >> 3, p & 7
        cur = d[byte] & (0xFF >> off)
        n = 0
        while cur == 0:
            n += 8 - off
            off = 0
            byte += 1
            cur = d[byte]
        lz = _LEADING_ZEROS[cur] - off
        self.pos = (byte << 3) + off + lz + 1
        return n + lz

    def align(self):
        self.pos = (self.pos + 7) & ~7


def _residual(br: _Bits, blocksize: int, order: int) -> List[int]:
    method = br.read(2)
    if method > 1:
        raise ValueError("reserved residual coding method")
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    porder = br.read(4)
    nparts = 1 << porder
    out: List[int] = []
    for part in range(nparts):
        n = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.read(pbits)
        if k == esc:
            nb = br.read(5)
            out.extend(br.read_signed(nb) for _ in range(n))
        else:
            unary, read = br.unary, br.read
            for _ in range(n):
                u = (unary() << k) | read(k)
                out.append((u >> 1) ^ -(u & 1))
    return out


_FIXED = ((), (1,), (2, -1), (3, -3, 1), (4, -6, 4, -1))


def _subframe(br: _Bits, blocksize: int, bps: int) -> List[int]:
    if br.read(1):
        raise ValueError("subframe padding bit set")
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.unary() + 1
        bps -= wasted
    if typ == 0:                                   # CONSTANT
        out = [br.read_signed(bps)] * blocksize
    elif typ == 1:                                 # VERBATIM
        out = [br.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:                           # FIXED predictor of order typ - 8
        order = typ - 8
        out = [br.read_signed(bps) for _ in range(order)]
        coefs = _FIXED[order]
        res = _residual(br, blocksize, order)
        for r in res:
            p = 0
            for j, c in enumerate(coefs):
                p += c * out[-1 - j]
            out.append(p + r)
    elif 32 <= typ <= 63:                          # LPC of order typ - 31
        order = typ - 31
        out = [br.read_signed(bps) for _ in range(order)]
        prec = br.read(4) + 1
        shift = br.read_signed(5)
        coefs = [br.read_signed(prec) for _ in range(order)]
        res = _residual(br, blocksize, order)
        for r in res:
            p = 0
            for j in range(order):
                p += coefs[j] * out[-1 - j]
            out.append((p >> shift) + r)
    else:
        raise ValueError(f"reserved subframe type {typ}")
    if wasted:
        out = [v << wasted for v in out]
    return out


_BLOCK = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}
_RATES = {1: 88200, 2: 176400, 3: 192000, 4: 8000, 5: 16000, 6: 22050, 7: 24000, 8: 32000, 9: 44100, 10: 48000, 11: 96000}
_BPS = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}


def _utf8_number(br: _Bits) -> int:
    b = br.read(8)
    if b < 0x80:
        return b
    n = 0
    while b & (0x80 >> n):
        n += 1
    v = b & (0x7F >> n)
    for _ in range(n - 1):
        v = (v << 6) | (br.read(8) & 0x3F)
    return v


def read_flac(path: str, verify: bool = True) -> Tuple[np.ndarray, int]:
    """-> (samples int32, shape (n,) for mono or (n, channels)), sample_rate)."""
    with open(path, "rb") as f:
        d = f.read()
    if d[:4] != b"fLaC":
        raise ValueError(f"{path}: not a FLAC stream")
    pos, info = 4, None
    while True:
        hdr = d[pos]
        ln = int.from_bytes(d[pos + 1:pos + 4], "big")
        if hdr & 0x7F == 0:
            body = d[pos + 4:pos + 4 + ln]
            x = int.from_bytes(body[10:18], "big")
            info = {"rate": x >> 44, "channels": ((x >> 41) & 7) + 1, "bps": ((x >> 36) & 31) + 1,
                    "total": x & ((1 << 36) - 1), "md5": body[18:34], "max_block": struct.unpack(">H", body[2:4])[0]}
        pos += 4 + ln
        if hdr & 0x80:
            break
    if info is None:
        raise ValueError(f"{path}: no STREAMINFO block")
    nch, total = info["channels"], info["total"]
    chans: List[List[int]] = [[] for _ in range(nch)]
    end = len(d)
    while pos + 2 <= end and (total == 0 or len(chans[0]) < total):
        br = _Bits(d, pos)
        if br.read(15) != 0x7FFC:                  # sync code 11111111111110 + reserved 0
            raise ValueError(f"{path}: lost frame sync at byte {pos}")
        br.read(1)                                 # blocking strategy (the coded number is not needed for decoding)
        bs_code, sr_code = br.read(4), br.read(4)
        ch_code, bps_code = br.read(4), br.read(3)
        br.read(1)
        _utf8_number(br)
        if bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        elif bs_code >= 8:
            blocksize = 256 << (bs_code - 8)
        elif bs_code in _BLOCK:
            blocksize = _BLOCK[bs_code]
        else:
            raise ValueError("reserved block size code")
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)                                 # CRC-8 of the header
        bps = _BPS.get(bps_code, info["bps"])
        if ch_code < 8:
            if ch_code + 1 != nch:
                raise ValueError("channel count changes mid-stream")
            subs = [_subframe(br, blocksize, bps) for _ in range(nch)]
        elif ch_code == 8:                         # left / side
            l, s = _subframe(br, blocksize, bps), _subframe(br, blocksize, bps + 1)
            subs = [l, [a - b for a, b in zip(l, s)]]
        elif ch_code == 9:                         # side / right
            s, r = _subframe(br, blocksize, bps + 1), _subframe(br, blocksize, bps)
            subs = [[a + b for a, b in zip(s, r)], r]
        elif ch_code == 10:                        # mid / side
            m, s = _subframe(br, blocksize, bps), _subframe(br, blocksize, bps + 1)
            l_, r_ = [], []
            for a, b in zip(m, s):
                a = (a << 1) | (b & 1)
                l_.append((a + b) >> 1)
                r_.append((a - b) >> 1)
            subs = [l_, r_]
        else:
            raise ValueError("reserved channel assignment")
        for c in range(nch):
            chans[c].extend(subs[c])
        br.align()
        br.read(16)                                # CRC-16 of the frame
        pos = br.pos >> 3
    pcm = np.array(chans, dtype=np.int64).T        # (n, channels)
    if total:
        pcm = pcm[:total]
    if verify and any(info["md5"]):
        width = (info["bps"] + 7) // 8
        raw = b"".join(int(v).to_bytes(width, "little", signed=True) for v in pcm.reshape(-1)) if width != 2 else \
            pcm.astype("<i2").tobytes()
        if hashlib.md5(raw).digest() != info["md5"]:
            raise ValueError(f"{path}: decoded PCM does not match the MD5 signature in STREAMINFO")
    out = pcm.astype(np.int32)
    return (out[:, 0] if nch == 1 else out), info["rate"]

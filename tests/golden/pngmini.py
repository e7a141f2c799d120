"""Minimal PNG / PNM / PAM readers used ONLY to turn the reference's test images into fixtures.

`read_png_as_iw_input` restates what the reference's PNG reader hands to the core library
(reference: src/imagew-png.c:136-155 colourspace, :231-256 bKGD, :295-400 pixel layout): palette images are
expanded to RGB(A) 8-bit, tRNS on gray/RGB becomes a full alpha channel (promoting <8-bit gray to 8-bit),
1/2/4-bit gray stays packed, 16-bit stays big-endian.  Non-interlaced files only (all the reference's test
images are).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


class Png:
    pass


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    if pa <= pb and pa <= pc:
        return a
    return b if pb <= pc else c


def read_png(path: str) -> Png:
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n", path
    pos = 8
    png = Png()
    png.plte = png.trns = png.bkgd = png.sbit = png.phys = None
    png.srgb = None
    png.gama = None
    idat = b""
    while pos < len(d):
        n, = struct.unpack(">I", d[pos:pos + 4])
        t = d[pos + 4:pos + 8]
        body = d[pos + 8:pos + 8 + n]
        pos += 12 + n
        if t == b"IHDR":
            png.width, png.height, png.depth, png.ctype, _, _, il = struct.unpack(">IIBBBBB", body)
            if il != 0:
                raise NotImplementedError("interlaced PNG not supported by this fixture reader")
        elif t == b"PLTE":
            png.plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif t == b"tRNS":
            png.trns = body
        elif t == b"bKGD":
            png.bkgd = body
        elif t == b"sBIT":
            png.sbit = body
        elif t == b"pHYs":
            png.phys = struct.unpack(">IIB", body)
        elif t == b"sRGB":
            png.srgb = body[0]
        elif t == b"gAMA":
            png.gama = struct.unpack(">I", body)[0]
        elif t == b"IDAT":
            idat += body
    raw = zlib.decompress(idat)
    nch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[png.ctype]
    bpp_bits = nch * png.depth
    bpr = (png.width * bpp_bits + 7) // 8
    fb = max(1, bpp_bits // 8)
    rows = np.zeros((png.height, bpr), np.uint8)
    prev = bytearray(bpr)
    p = 0
    for y in range(png.height):
        ft = raw[p]
        line = bytearray(raw[p + 1:p + 1 + bpr])
        p += 1 + bpr
        if ft == 1:
            for i in range(fb, bpr):
                line[i] = (line[i] + line[i - fb]) & 255
        elif ft == 2:
            for i in range(bpr):
                line[i] = (line[i] + prev[i]) & 255
        elif ft == 3:
            for i in range(bpr):
                a = line[i - fb] if i >= fb else 0
                line[i] = (line[i] + ((a + prev[i]) >> 1)) & 255
        elif ft == 4:
            for i in range(bpr):
                a = line[i - fb] if i >= fb else 0
                c = prev[i - fb] if i >= fb else 0
                line[i] = (line[i] + _paeth(a, prev[i], c)) & 255
        rows[y] = np.frombuffer(bytes(line), np.uint8)
        prev = line
    png.rows = rows
    png.nch = nch
    return png


def _unpack_samples(png: Png) -> np.ndarray:
    """(H, W*nch) integer samples at native depth."""
    if png.depth == 8:
        return png.rows.astype(np.uint32)
    if png.depth == 16:
        r = png.rows.reshape(png.height, -1, 2).astype(np.uint32)
        return r[:, :, 0] * 256 + r[:, :, 1]
    bits = np.unpackbits(png.rows, axis=1)
    n = png.width * png.nch
    bits = bits[:, :n * png.depth].reshape(png.height, n, png.depth).astype(np.uint32)
    w = (1 << np.arange(png.depth - 1, -1, -1)).astype(np.uint32)
    return (bits * w).sum(axis=2)


IMGTYPE_GRAY, IMGTYPE_GRAYA, IMGTYPE_RGB, IMGTYPE_RGBA = 0x0001, 0x0101, 0x0010, 0x0110


def _pack_rows(samples: np.ndarray, depth: int) -> np.ndarray:
    """(H, N) integer samples -> (H, bpr) bytes, big-endian for 16-bit."""
    h = samples.shape[0]
    if depth == 8:
        return samples.astype(np.uint8)
    if depth == 16:
        out = np.zeros((h, samples.shape[1], 2), np.uint8)
        out[:, :, 0] = samples >> 8
        out[:, :, 1] = samples & 255
        return out.reshape(h, -1)
    raise ValueError(depth)


def read_png_as_iw_input(path: str) -> dict:
    """What `iw_read_png_file` would pass to iw_set_input_image + friends."""
    png = read_png(path)
    s = _unpack_samples(png)
    h, w = png.height, png.width
    info = {"width": w, "height": h, "native_grayscale": int(not (png.ctype & 2)), "bkgd_label": None,
            "sbit": png.sbit is not None, "phys": png.phys}
    depth = png.depth
    if png.ctype == 3:
        pal = np.zeros((256, 4), np.uint32)
        pal[:, 3] = 255
        pal[:len(png.plte), :3] = png.plte
        if png.trns is not None:
            t = np.frombuffer(png.trns, np.uint8)
            pal[:len(t), 3] = t
            px = pal[s]                     # H, W, 4
            imgtype = IMGTYPE_RGBA
        else:
            px = pal[s][:, :, :3]
            imgtype = IMGTYPE_RGB
        depth = 8
        rows = _pack_rows(px.reshape(h, -1), 8)
    elif png.trns is not None and png.ctype in (0, 2):
        # tRNS colour key -> alpha channel; <8-bit gray is expanded to 8 bits (libpng tRNS_to_alpha)
        if png.ctype == 0:
            key = struct.unpack(">H", png.trns[:2])[0]
            alpha = np.where(s == key, 0, (1 << max(depth, 8)) - 1 if depth == 16 else 255)
            g = s
            if depth < 8:
                g = s * (255 // ((1 << depth) - 1))
                depth = 8
            px = np.stack([g, alpha], axis=2)
            imgtype = IMGTYPE_GRAYA
        else:
            key = struct.unpack(">HHH", png.trns[:6])
            rgb = s.reshape(h, w, 3)
            m = (rgb[:, :, 0] == key[0]) & (rgb[:, :, 1] == key[1]) & (rgb[:, :, 2] == key[2])
            alpha = np.where(m, 0, (1 << depth) - 1)
            px = np.concatenate([rgb, alpha[:, :, None]], axis=2)
            imgtype = IMGTYPE_RGBA
        rows = _pack_rows(px.reshape(h, -1), depth)
    else:
        imgtype = {0: IMGTYPE_GRAY, 2: IMGTYPE_RGB, 4: IMGTYPE_GRAYA, 6: IMGTYPE_RGBA}[png.ctype]
        rows = png.rows.copy()
    info.update(imgtype=imgtype, bit_depth=depth, rows=rows)
    # colourspace: sRGB chunk -> sRGB; gAMA -> gamma 1/g; default sRGB
    if png.srgb is not None or png.gama is None:
        info["cs"] = ("srgb", 0.0)
    else:
        g = int(0.5 + 100000.0 * (png.gama / 100000.0)) / 100000.0
        info["cs"] = ("gamma", 1.0 / g)
    if png.bkgd is not None:
        # NB the reference divides by the ORIGINAL bit depth's max (rctx->bit_depth is updated after
        # png_read_update_info, which for palette images makes it 8)
        if png.ctype == 3:
            idx = png.bkgd[0]
            mx = 255.0
            rgb = [float(x) for x in png.plte[idx]] if idx < len(png.plte) else [0.0, 0.0, 0.0]
            info["bkgd_label"] = [rgb[0] / mx, rgb[1] / mx, rgb[2] / mx]
        elif png.ctype in (0, 4):
            mx = float((1 << depth) - 1)
            gval = struct.unpack(">H", png.bkgd[:2])[0]
            info["bkgd_label"] = [gval / mx] * 3
        else:
            mx = float((1 << depth) - 1)
            r, g, b = struct.unpack(">HHH", png.bkgd[:6])
            info["bkgd_label"] = [r / mx, g / mx, b / mx]
    return info


def png_to_canonical(path: str):
    """Decode any (non-interlaced) PNG to (depth, RGBA uint16 array HxWx4) in native sample scale
    (8-bit files -> 0..255, 16-bit -> 0..65535; low-depth gray scaled up to 8)."""
    png = read_png(path)
    s = _unpack_samples(png)
    h, w = png.height, png.width
    depth = 16 if png.depth == 16 else 8
    mx = 65535 if depth == 16 else 255
    out = np.zeros((h, w, 4), np.uint32)
    out[:, :, 3] = mx
    if png.ctype == 3:
        pal = np.zeros((256, 4), np.uint32)
        pal[:, 3] = 255
        pal[:len(png.plte), :3] = png.plte
        if png.trns is not None:
            t = np.frombuffer(png.trns, np.uint8)
            pal[:len(t), 3] = t
        out = pal[s]
    elif png.ctype == 0:
        g = s if png.depth >= 8 else s * (255 // ((1 << png.depth) - 1))
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = g
        if png.trns is not None:
            key = struct.unpack(">H", png.trns[:2])[0]
            out[:, :, 3] = np.where(s == key, 0, mx)
    elif png.ctype == 2:
        rgb = s.reshape(h, w, 3)
        out[:, :, :3] = rgb
        if png.trns is not None:
            key = struct.unpack(">HHH", png.trns[:6])
            m = (rgb[:, :, 0] == key[0]) & (rgb[:, :, 1] == key[1]) & (rgb[:, :, 2] == key[2])
            out[:, :, 3] = np.where(m, 0, mx)
    elif png.ctype == 4:
        ga = s.reshape(h, w, 2)
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = ga[:, :, 0]
        out[:, :, 3] = ga[:, :, 1]
    else:
        out = s.reshape(h, w, 4)
    return depth, out.astype(np.uint16), png


def pnm_to_canonical(path: str):
    """Binary PGM/PPM/PAM -> (maxval, RGBA uint16 HxWx4 in raw code units)."""
    d = open(path, "rb").read()
    if d[:2] == b"P7":
        hdr_end = d.index(b"ENDHDR\n") + 7
        fields = {}
        for line in d[3:hdr_end].decode("ascii").splitlines():
            parts = line.split()
            if len(parts) >= 2:
                fields[parts[0]] = parts[1]
        w, h, nch, mx = int(fields["WIDTH"]), int(fields["HEIGHT"]), int(fields["DEPTH"]), int(fields["MAXVAL"])
        body = d[hdr_end:]
    else:
        kind = d[:2]
        if kind not in (b"P4", b"P5", b"P6"):
            raise NotImplementedError("only binary PBM/PGM/PPM/PAM are modelled")
        ntok = 2 if kind == b"P4" else 3
        toks = []
        pos = 2
        while len(toks) < ntok:
            if pos >= len(d):
                raise ValueError("truncated PNM header")
            ch = d[pos:pos + 1]
            if ch.isspace():
                pos += 1
            elif ch == b"#":
                while pos < len(d) and d[pos:pos + 1] != b"\n":
                    pos += 1
            else:
                st = pos
                while pos < len(d) and not d[pos:pos + 1].isspace():
                    pos += 1
                toks.append(int(d[st:pos]))
        pos += 1
        if kind == b"P4":
            w, h = toks
            bpr = (w + 7) // 8
            bits = np.unpackbits(np.frombuffer(d[pos:pos + bpr * h], np.uint8).reshape(h, bpr), axis=1)[:, :w]
            out = np.zeros((h, w, 4), np.uint32)
            out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = 1 - bits   # PBM: 1 = black
            out[:, :, 3] = 1
            return 1, out.astype(np.uint16)
        w, h, mx = toks
        nch = {b"P5": 1, b"P6": 3}[kind]
        body = d[pos:]
    if mx > 255:
        s = np.frombuffer(body[:w * h * nch * 2], np.uint8).reshape(h, w, nch, 2).astype(np.uint32)
        s = s[..., 0] * 256 + s[..., 1]
    else:
        s = np.frombuffer(body[:w * h * nch], np.uint8).reshape(h, w, nch).astype(np.uint32)
    out = np.zeros((h, w, 4), np.uint32)
    out[:, :, 3] = mx
    if nch == 1:
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = s[:, :, 0]
    elif nch == 2:
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = s[:, :, 0]
        out[:, :, 3] = s[:, :, 1]
    elif nch == 3:
        out[:, :, :3] = s
    else:
        out = s
    return mx, out.astype(np.uint16)

"""CPU oracle for the compressibility scorer.  TEST INFRASTRUCTURE ONLY.

The reference scores `1 - clip(len(JPEG bytes)/3000, 0, 1)` with the bytes produced by
`PIL.Image.save(format='JPEG', quality=80)` (edm/scorers.py:207-244), i.e. by libjpeg(-turbo), a
third-party dependency that is not part of /root/reference (environment.yml: pillow>=8.3.1; this
image: Pillow + libjpeg-turbo).  This module restates the published baseline-JPEG algorithm as
libjpeg implements it (jccolor.c rgb_ycc_convert, jcsample.c h2v2_downsample, jfdctint.c
jpeg_fdct_islow, jcdctmgr.c quantisation, jchuff.c encode_one_block + byte stuffing) in numpy integer
arithmetic, and is pinned against PIL's real byte counts (tests/test_jpeg_oracle.py)."""
import io

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63])


def parse_jpeg_tables(data: bytes):
    """Quantisation tables (natural order), Huffman (bits, huffval) tables and the entropy-data offset of a
    baseline JPEG file."""
    qt, ht = {}, {}
    i = 2
    while i < len(data):
        assert data[i] == 0xFF
        marker = data[i + 1]
        seglen = (data[i + 2] << 8) | data[i + 3]
        body = data[i + 4:i + 2 + seglen]
        if marker == 0xDB:
            j = 0
            while j < len(body):
                tq = body[j] & 15
                tab = np.zeros(64, dtype=np.int32)
                tab[ZIGZAG] = np.frombuffer(body[j + 1:j + 65], dtype=np.uint8)
                qt[tq] = tab
                j += 65
        elif marker == 0xC4:
            j = 0
            while j < len(body):
                tc, th = body[j] >> 4, body[j] & 15
                bits = list(body[j + 1:j + 17])
                n = sum(bits)
                ht[(tc, th)] = (bits, list(body[j + 17:j + 17 + n]))
                j += 17 + n
        elif marker == 0xDA:
            return qt, ht, i + 2 + seglen
        i += 2 + seglen
    raise ValueError('no SOS marker')


def code_lengths(bits, vals):
    """symbol -> Huffman code length (jchuff.c jpeg_make_c_derived_tbl) and symbol -> code."""
    size, code_of, code, k = np.zeros(256, dtype=np.int32), np.zeros(256, dtype=np.int64), 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            size[vals[k]] = length
            code_of[vals[k]] = code
            code += 1
            k += 1
        code <<= 1
    return size, code_of


def pil_tables(width, height, quality):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.zeros((height, width, 3), dtype=np.uint8)).save(buf, format='JPEG', quality=quality)
    data = buf.getvalue()
    qt, ht, sos_end = parse_jpeg_tables(data)
    return qt, ht, sos_end


def pil_size(img_hwc: np.ndarray, quality=80) -> int:
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img_hwc).save(buf, format='JPEG', quality=quality)
    return len(buf.getvalue())


def _fix(x):
    return int(x * 65536 + 0.5)


def rgb_to_ycc(img):
    """jccolor.c rgb_ycc_convert (SCALEBITS 16)."""
    r, g, b = [img[..., i].astype(np.int64) for i in range(3)]
    half, off = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def h2v2_downsample(p):
    """jcsample.c h2v2_downsample: bias alternates 1,2,1,2 along each output row."""
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = np.tile(np.array([1, 2]), s.shape[1] // 2 + 1)[:s.shape[1]]
    return (s + bias[None, :]) >> 2


def fdct_islow(block):
    """jfdctint.c jpeg_fdct_islow on an 8x8 block of (sample - 128); output scaled by 8."""
    CB, P1 = 13, 2
    F = lambda x: int(x * (1 << CB) + 0.5)
    c = dict(f0_298=F(0.298631336), f0_390=F(0.390180644), f0_541=F(0.541196100), f0_765=F(0.765366865),
             f0_899=F(0.899976223), f1_175=F(1.175875602), f1_501=F(1.501321110), f1_847=F(1.847759065),
             f1_961=F(1.961570560), f2_053=F(2.053119869), f2_562=F(2.562915447), f3_072=F(3.072711026))
    d = block.astype(np.int64).copy()

    def desc(x, n):
        return (x + (1 << (n - 1))) >> n

    def pass_(v, first):
        t0, t7 = v[0] + v[7], v[0] - v[7]
        t1, t6 = v[1] + v[6], v[1] - v[6]
        t2, t5 = v[2] + v[5], v[2] - v[5]
        t3, t4 = v[3] + v[4], v[3] - v[4]
        t10, t13 = t0 + t3, t0 - t3
        t11, t12 = t1 + t2, t1 - t2
        o = [0] * 8
        if first:
            o[0] = (t10 + t11) << P1
            o[4] = (t10 - t11) << P1
        else:
            o[0] = desc(t10 + t11, P1)
            o[4] = desc(t10 - t11, P1)
        z1 = (t12 + t13) * c['f0_541']
        sh = CB - P1 if first else CB + P1
        o[2] = desc(z1 + t13 * c['f0_765'], sh)
        o[6] = desc(z1 + t12 * (-c['f1_847']), sh)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * c['f1_175']
        t4, t5, t6, t7 = t4 * c['f0_298'], t5 * c['f2_053'], t6 * c['f3_072'], t7 * c['f1_501']
        z1, z2, z3, z4 = z1 * (-c['f0_899']), z2 * (-c['f2_562']), z3 * (-c['f1_961']), z4 * (-c['f0_390'])
        z3, z4 = z3 + z5, z4 + z5
        o[7] = desc(t4 + z1 + z3, sh)
        o[5] = desc(t5 + z2 + z4, sh)
        o[3] = desc(t6 + z2 + z3, sh)
        o[1] = desc(t7 + z1 + z4, sh)
        return o
    for r in range(8):
        d[r, :] = pass_([int(x) for x in d[r, :]], True)
    for col in range(8):
        d[:, col] = pass_([int(x) for x in d[:, col]], False)
    return d


def quantize(coef, qtab):
    """jcdctmgr.c: divisor = q << 3 (islow), round half away from zero."""
    q = (qtab.reshape(8, 8).astype(np.int64)) << 3
    a = np.abs(coef) + (q >> 1)
    return np.sign(coef) * (a // q)


def _nbits(v):
    v = abs(int(v))
    n = 0
    while v:
        n += 1
        v >>= 1
    return n


def encode_image(img_hwc: np.ndarray, quality=80):
    """Returns (file_size, entropy_bits) of the baseline 4:2:0 JPEG libjpeg would write."""
    H, W, _ = img_hwc.shape
    assert H % 16 == 0 and W % 16 == 0
    qt, ht, sos_end = pil_tables(W, H, quality)
    dc_size = [code_lengths(*ht[(0, 0)]), code_lengths(*ht[(0, 1)])]
    ac_size = [code_lengths(*ht[(1, 0)]), code_lengths(*ht[(1, 1)])]
    y, cb, cr = rgb_to_ycc(img_hwc)
    planes = [y - 128, h2v2_downsample(cb) - 128, h2v2_downsample(cr) - 128]
    bits = []          # list of (value, nbits)
    pred = [0, 0, 0]

    def put(code, n):
        bits.append((int(code), int(n)))

    def encode_block(comp, by, bx):
        tabs = 0 if comp == 0 else 1
        blk = planes[comp][by * 8:by * 8 + 8, bx * 8:bx * 8 + 8]
        zz = quantize(fdct_islow(blk), qt[tabs]).reshape(64)[ZIGZAG]
        diff = int(zz[0]) - pred[comp]
        pred[comp] = int(zz[0])
        n = _nbits(diff)
        put(dc_size[tabs][1][n], dc_size[tabs][0][n])
        if n:
            put((diff if diff >= 0 else diff - 1) & ((1 << n) - 1), n)
        run = 0
        for k in range(1, 64):
            v = int(zz[k])
            if v == 0:
                run += 1
                continue
            while run > 15:
                put(ac_size[tabs][1][0xF0], ac_size[tabs][0][0xF0])
                run -= 16
            n = _nbits(v)
            put(ac_size[tabs][1][(run << 4) | n], ac_size[tabs][0][(run << 4) | n])
            put((v if v >= 0 else v - 1) & ((1 << n) - 1), n)
            run = 0
        if run:
            put(ac_size[tabs][1][0], ac_size[tabs][0][0])
    for my in range(H // 16):
        for mx in range(W // 16):
            for dy in range(2):
                for dx in range(2):
                    encode_block(0, my * 2 + dy, mx * 2 + dx)
            encode_block(1, my, mx)
            encode_block(2, my, mx)
    total_bits = sum(n for _, n in bits)
    # materialise the stream to count 0xFF bytes (each is followed by a stuffed 0x00); pad with 1-bits
    acc, nacc, nbytes, nff = 0, 0, 0, 0
    for code, n in bits:
        acc = (acc << n) | code
        nacc += n
        while nacc >= 8:
            byte = (acc >> (nacc - 8)) & 0xFF
            nacc -= 8
            nbytes += 1
            nff += byte == 0xFF
        acc &= (1 << nacc) - 1
    if nacc:
        byte = ((acc << (8 - nacc)) | ((1 << (8 - nacc)) - 1)) & 0xFF
        nbytes += 1
        nff += byte == 0xFF
    return sos_end + nbytes + nff + 2, total_bits


def compressibility_score(img_hwc, quality=80, min_size=0, max_size=3000, size=None):
    """edm/scorers.py:243: 1 - clip((size - min)/(max - min), 0, 1)."""
    size = encode_image(img_hwc, quality)[0] if size is None else size
    return 1.0 - min(1.0, max(0.0, (size - min_size) / (max_size - min_size)))

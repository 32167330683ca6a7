"""A small pure-Python BAM reader / writer (struct + zlib), independent of csrc/host: test infrastructure that
cross-checks the C++ decoder and writers."""
import struct
import zlib

NIB = "=ACMGRSVTWYHKDBN"
REF_OPS = (0, 2, 3, 7, 8)


def bgzf_read(path):
    data = open(path, "rb").read()
    out, p = [], 0
    while p < len(data):
        assert data[p : p + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", data, p + 10)[0]
        x, bsize = p + 12, None
        while x < p + 12 + xlen:
            si1, si2, slen = data[x], data[x + 1], struct.unpack_from("<H", data, x + 2)[0]
            if (si1, si2) == (66, 67):
                bsize = struct.unpack_from("<H", data, x + 4)[0]
            x += 4 + slen
        payload = data[p + 12 + xlen : p + bsize + 1 - 8]
        raw = zlib.decompress(payload, -15)
        crc, isize = struct.unpack_from("<II", data, p + bsize + 1 - 8)
        assert zlib.crc32(raw) == crc and len(raw) == isize
        out.append(raw)
        p += bsize + 1
    assert out and out[-1] == b"", "BGZF EOF marker missing"
    return b"".join(out)


def bgzf_write(path, raw):
    with open(path, "wb") as f:
        for off in list(range(0, len(raw), 0xFF00)) + [None]:
            chunk = b"" if off is None else raw[off : off + 0xFF00]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def parse_aux(b):
    tags, order, p = {}, [], 0
    sizes = {"A": ("c", 1), "c": ("b", 1), "C": ("B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4), "f": ("<f", 4)}
    while p < len(b):
        tag, ty = b[p : p + 2].decode(), chr(b[p + 2])
        p += 3
        if ty in sizes:
            fmt, n = sizes[ty]
            val = struct.unpack_from(fmt, b, p)[0]
            p += n
        elif ty in "ZH":
            e = b.index(b"\0", p)
            val = b[p:e].decode()
            p = e + 1
        elif ty == "B":
            sub, cnt = chr(b[p]), struct.unpack_from("<I", b, p + 1)[0]
            fmt, n = sizes[sub]
            val = [struct.unpack_from(fmt, b, p + 5 + k * n)[0] for k in range(cnt)]
            p += 5 + cnt * n
        else:
            raise ValueError(ty)
        tags[tag] = (ty, val)
        order.append(tag)
    return tags, order


def read_bam(path):
    raw = bgzf_read(path)
    assert raw[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", raw, 4)[0]
    text = raw[8 : 8 + l_text]
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", raw, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<I", raw, p)[0]
        name = raw[p + 4 : p + 4 + ln - 1].decode()
        refs.append((name, struct.unpack_from("<I", raw, p + 4 + ln)[0]))
        p += 8 + ln
    header_bytes = raw[:p]
    recs = []
    while p < len(raw):
        bs = struct.unpack_from("<I", raw, p)[0]
        b = raw[p + 4 : p + 4 + bs]
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHI", b, 0)
        q = 32
        qname = b[q : q + l_name - 1].decode()
        q += l_name
        cigar = list(struct.unpack_from(f"<{n_cig}I", b, q))
        q += 4 * n_cig
        seq = "".join(NIB[(b[q + (k >> 1)] >> (0 if k & 1 else 4)) & 15] for k in range(l_seq))
        q += (l_seq + 1) // 2
        qual = bytes(b[q : q + l_seq])
        q += l_seq
        tags, order = parse_aux(b[q:])
        rlen = 0 if flag & 4 else sum(c >> 4 for c in cigar if (c & 15) in REF_OPS)
        recs.append({"raw": bytes(b), "core_and_data": bytes(b[:q]), "aux": bytes(b[q:]), "tid": tid, "pos": pos, "end": pos + (rlen or 1), "mapq": mapq, "flag": flag, "qname": qname,
                     "cigar": cigar, "seq": seq, "qual": qual, "tags": tags, "tag_order": order})
        p += 4 + bs
    return {"text": text, "refs": refs, "header_bytes": header_bytes, "records": recs}


def make_record(tid, pos, qname, cigar, flag=0, mapq=60, aux=b""):
    """cigar: [(op, len)]; bases all 'A', qualities 30."""
    l_seq = sum(ln for op, ln in cigar if op in (0, 1, 4, 7, 8))
    body = struct.pack("<iiBBHHHIiii", tid, pos, len(qname) + 1, mapq, 4680, len(cigar), flag, l_seq, -1, -1, 0)
    body += qname.encode() + b"\0" + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar) + b"\x11" * ((l_seq + 1) // 2) + bytes([30]) * l_seq + aux
    return struct.pack("<I", len(body)) + body


def write_bam(path, refs, records, text=None):
    if text is None:
        text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    raw = b"BAM\1" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", len(refs))
    for n, l in refs:
        raw += struct.pack("<I", len(n) + 1) + n.encode() + b"\0" + struct.pack("<I", l)
    bgzf_write(path, raw + b"".join(records))

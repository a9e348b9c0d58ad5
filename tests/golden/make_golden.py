#!/usr/bin/env python3
"""Builds tests/golden/reference_vectors.{json,bin} from the reference's own golden vectors.

Run in the build container only (needs /root/reference); the GPU box uses the committed outputs.
Sources (gendx/lzma-rs @ 1f14478):
  * fixtures tests/files/*.lzma, *.xz with their plaintexts (tests/lzma.rs:171-195, tests/xz.rs:63-83,112-120)
  * inline known-answer vectors of tests/lzma.rs:198-234, tests/xz.rs:86-109
  * expected error strings of tests/xz.rs:123-146, tests/lzma.rs:135-143,306-336, src/decode/stream.rs:376-388,461-467
Compressed inputs go into one binary blob; plaintexts are pinned by sha256 + length (small ones also inline).
"""
import hashlib
import json
import os

REF = "/root/reference/tests/files"
HERE = os.path.dirname(os.path.abspath(__file__))

FIXTURES = [  # (name, format, compressed file, plaintext file)
    ("foo.txt.lzma", "lzma", "foo.txt.lzma", "foo.txt"),
    ("hugedict.txt.lzma", "lzma", "hugedict.txt.lzma", "foo.txt"),
    ("range-coder-edge-case.lzma", "lzma", "range-coder-edge-case.lzma", "range-coder-edge-case"),
    ("hello.txt.lzma", "lzma", "hello.txt.lzma", "hello.txt"),
    ("empty.txt.lzma", "lzma", "empty.txt.lzma", "empty.txt"),
    ("foo.txt.xz", "xz", "foo.txt.xz", "foo.txt"),
    ("good-1-lzma2-1.xz", "xz", "good-1-lzma2-1.xz", "good-1-lzma2-1"),
    ("good-1-lzma2-2.xz", "xz", "good-1-lzma2-2.xz", "good-1-lzma2-2"),
    ("good-1-lzma2-3.xz", "xz", "good-1-lzma2-3.xz", "good-1-lzma2-3"),
    ("good-1-lzma2-4.xz", "xz", "good-1-lzma2-4.xz", "good-1-lzma2-4"),
    ("hello.txt.xz", "xz", "hello.txt.xz", "hello.txt"),
    ("empty.txt.xz", "xz", "empty.txt.xz", "empty.txt"),
    ("block-check-crc32.txt.xz", "xz", "block-check-crc32.txt.xz", "block-check-crc32.txt"),
]

INLINE = [  # (name, format, compressed bytes, plaintext)
    ("inline-empty.lzma", "lzma",  # tests/lzma.rs:198-208
     b"\x5d\x00\x00\x80\x00\xff\xff\xff\xff\xff\xff\xff\xff\x00\x83\xff\xfb\xff\xff\xc0\x00\x00\x00", b""),
    ("inline-hello.lzma", "lzma",  # tests/lzma.rs:210-221
     b"\x5d\x00\x00\x80\x00\xff\xff\xff\xff\xff\xff\xff\xff\x00\x24\x19\x49\x98\x6f\x10\x19\xc6\xd7\x31\xeb\x36"
     b"\x50\xb2\x98\x48\xff\xfe\xa5\xb0\x00", b"Hello world\x0a"),
    ("inline-hello-hugedict.lzma", "lzma",  # tests/lzma.rs:223-234
     b"\x5d\x7f\x7f\x7f\x7f\xff\xff\xff\xff\xff\xff\xff\xff\x00\x24\x19\x49\x98\x6f\x10\x19\xc6\xd7\x31\xeb\x36"
     b"\x50\xb2\x98\x48\xff\xfe\xa5\xb0\x00", b"Hello world\x0a"),
    ("inline-empty.xz", "xz",  # tests/xz.rs:86-95
     b"\xfd\x37\x7a\x58\x5a\x00\x00\x04\xe6\xd6\xb4\x46\x00\x00\x00\x00\x1c\xdf\x44\x21\x1f\xb6\xf3\x7d\x01\x00"
     b"\x00\x00\x00\x04\x59\x5a", b""),
    ("inline-hello.xz", "xz",  # tests/xz.rs:97-109
     b"\xfd\x37\x7a\x58\x5a\x00\x00\x04\xe6\xd6\xb4\x46\x02\x00\x21\x01\x16\x00\x00\x00\x74\x2f\xe5\xa3\x01\x00"
     b"\x0b\x48\x65\x6c\x6c\x6f\x20\x77\x6f\x72\x6c\x64\x0a\x00\xca\xec\x49\x05\x66\x3f\x67\x98\x00\x01\x24\x0c"
     b"\xa6\x18\xd8\xd8\x1f\xb6\xf3\x7d\x01\x00\x00\x00\x00\x04\x59\x5a", b"Hello world\x0a"),
]


def main():
    blob = bytearray()
    vectors = []

    def add(name, fmt, comp, plain, **extra):
        v = {"name": name, "format": fmt, "offset": len(blob), "length": len(comp),
             "plain_len": len(plain), "plain_sha256": hashlib.sha256(plain).hexdigest()}
        if len(plain) <= 4096:
            v["plain_hex"] = plain.hex()
        v.update(extra)
        blob.extend(comp)
        vectors.append(v)

    for name, fmt, cf, pf in FIXTURES:
        add(name, fmt, open(os.path.join(REF, cf), "rb").read(), open(os.path.join(REF, pf), "rb").read())
    for name, fmt, comp, plain in INLINE:
        add(name, fmt, comp, plain)

    # error vectors (expected Display strings of error::Error)
    bc = bytearray(open(os.path.join(REF, "block-check-crc32.txt.xz"), "rb").read())
    bc[0x54:0x58] = bytes([0x67, 0x45, 0x23, 0x01])  # tests/xz.rs:128-135
    plain = open(os.path.join(REF, "block-check-crc32.txt"), "rb").read()
    add("corrupt-footer.xz", "xz", bytes(bc), plain,  # block is valid, so its output reaches the sink first
        error="xz error: Invalid footer CRC32: expected 0x01234567 but got 0x8b0d303e", error_match="exact")
    add("empty-input.lzma", "lzma", b"", b"", error="header too short", error_match="prefix")  # tests/lzma.rs:135-143
    add("bad-props.lzma", "lzma", b"\xff" * 32, b"",  # src/decode/stream.rs:376-388
        error="LZMA header invalid properties: 255 must be < 225", error_match="contains")
    add("garbage.lzma", "lzma", b"corrupted bytes here corrupted bytes here", b"",  # stream.rs:461-467
        error="beyond output size", error_match="contains")

    with open(os.path.join(HERE, "reference_vectors.bin"), "wb") as f:
        f.write(blob)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump({"source": "gendx/lzma-rs @ 1f14478 tests/files + inline test vectors", "vectors": vectors}, f, indent=1)
    print(f"{len(vectors)} vectors, {len(blob)} bytes")


if __name__ == "__main__":
    main()

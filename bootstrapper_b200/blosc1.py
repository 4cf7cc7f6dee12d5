"""Blosc-1 frames as zarr / numcodecs write them (`Blosc(cname='lz4', clevel=5, shuffle=SHUFFLE)` is the zarr-v2 default and
what the reference's predictions arrive in).  `numcodecs` / `blosc` are not in this image; the container is parsed here
(c-blosc README_HEADER / blosc_d: 16-byte header, block starts, per-block splits, byte shuffle) and the inner codecs come
from what IS installed: LZ4 block and Zstandard through pyarrow, zlib from the standard library.  BloscLZ and Snappy
payloads raise.  `encode` writes frames of the same layout (LZ4 / zlib / zstd), used for outputs and tests.
"""
import struct
import zlib

import numpy as np

_CODECS = {0: "blosclz", 1: "lz4", 2: "snappy", 3: "zlib", 4: "zstd"}
_FORMAT = {"lz4": 1, "lz4hc": 1, "zlib": 3, "zstd": 4}
MAX_SPLITS, MIN_BUFFERSIZE = 16, 128


def _inner_decompress(codec, buf, n):
    if codec == "zlib":
        return zlib.decompress(buf)
    if codec in ("lz4", "zstd"):
        import pyarrow as pa
        return pa.decompress(buf, decompressed_size=n, codec="lz4_raw" if codec == "lz4" else "zstd").to_pybytes()
    raise ValueError(f"blosc inner codec {codec!r} is not available here (lz4, zstd, zlib are)")


def _inner_compress(codec, buf):
    if codec == "zlib":
        return zlib.compress(buf, 5)
    import pyarrow as pa
    return pa.compress(buf, codec="lz4_raw" if codec == "lz4" else "zstd", asbytes=True)


def _unshuffle(block, typesize):
    n = len(block) // typesize
    a = np.frombuffer(block, dtype=np.uint8)
    out = np.empty(len(block), dtype=np.uint8)
    out[:n * typesize] = a[:n * typesize].reshape(typesize, n).T.ravel()
    out[n * typesize:] = a[n * typesize:]
    return out.tobytes()


def _shuffle(block, typesize):
    n = len(block) // typesize
    a = np.frombuffer(block, dtype=np.uint8)
    out = np.empty(len(block), dtype=np.uint8)
    out[:n * typesize] = a[:n * typesize].reshape(n, typesize).T.ravel()
    out[n * typesize:] = a[n * typesize:]
    return out.tobytes()


def decode(buf):
    buf = bytes(buf)
    if len(buf) < 16:
        raise ValueError("not a blosc frame")
    version, versionlz, flags, typesize = buf[0], buf[1], buf[2], buf[3]
    nbytes, blocksize, cbytes = struct.unpack_from("<III", buf, 4)
    if flags & 0x04:
        raise ValueError("blosc bit-shuffle is not supported")
    if flags & 0x02:                                  # memcpyed
        return buf[16:16 + nbytes]
    codec = _CODECS.get(flags >> 5)
    doshuffle = bool(flags & 0x01) and typesize > 1
    dont_split = bool(flags & 0x10)
    nblocks = (nbytes + blocksize - 1) // blocksize if blocksize else 0
    bstarts = struct.unpack_from(f"<{nblocks}i", buf, 16)
    out = bytearray()
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        nsplits = typesize if (not dont_split and typesize <= MAX_SPLITS and blocksize // typesize >= MIN_BUFFERSIZE and not leftover) else 1
        neblock = bsize // nsplits
        pos = bstarts[b]
        parts = []
        for _ in range(nsplits):
            (cb,) = struct.unpack_from("<i", buf, pos)
            pos += 4
            chunk = buf[pos:pos + cb]
            pos += cb
            parts.append(chunk if cb == neblock else _inner_decompress(codec, chunk, neblock))
        block = b"".join(parts)
        if len(block) != bsize:
            raise ValueError("blosc block decodes to the wrong size")
        out += _unshuffle(block, typesize) if doshuffle else block
    return bytes(out)


def encode(data, typesize=1, cname="lz4", shuffle=1, blocksize=0):
    data = bytes(data)
    nbytes = len(data)
    codec = {"lz4hc": "lz4"}.get(cname, cname)
    if codec not in ("lz4", "zlib", "zstd"):
        raise ValueError(f"cannot write blosc frames with {cname!r}")
    blocksize = blocksize or min(max(nbytes, 1), 1 << 18)
    if nbytes < MIN_BUFFERSIZE:                       # tiny buffers are stored
        return struct.pack("<BBBBIII", 2, 1, 0x02 | (shuffle & 1), typesize, nbytes, max(nbytes, 1), 16 + nbytes) + data
    doshuffle = bool(shuffle & 1) and typesize > 1
    nblocks = (nbytes + blocksize - 1) // blocksize
    flags = (shuffle & 1) | (_FORMAT[cname] << 5)
    body = bytearray()
    bstarts = []
    base = 16 + 4 * nblocks
    for b in range(nblocks):
        block = data[b * blocksize:(b + 1) * blocksize]
        bsize = len(block)
        leftover = bsize != blocksize
        if doshuffle:
            block = _shuffle(block, typesize)
        nsplits = typesize if (typesize <= MAX_SPLITS and blocksize // typesize >= MIN_BUFFERSIZE and not leftover) else 1
        neblock = bsize // nsplits
        bstarts.append(base + len(body))
        for j in range(nsplits):
            part = block[j * neblock:(j + 1) * neblock]
            comp = _inner_compress(codec, part)
            if len(comp) >= neblock:
                comp = part                           # incompressible split: stored, marked by cbytes == neblock
            body += struct.pack("<i", len(comp)) + comp
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, base + len(body))
    return head + struct.pack(f"<{nblocks}i", *bstarts) + bytes(body)

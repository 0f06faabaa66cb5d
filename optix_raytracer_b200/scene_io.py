"""Model ingest and image output either side of the path that need no GPU: imgui_test's NBT model files and
the EXR branch of sutil::saveImage.  numpy + the standard library only.

* `load_nbt` — SDK/imgui_test/triangle_gas.cpp:16-76.  The reader library (awegsche/nbt@main, fetched by SDK/ext/CMakeLists.txt:12-17) is
  not vendored in the reference tree; the container format is the published Named Binary Tag one (big-endian, tag id / name / payload,
  optionally gzip- or zlib-wrapped), restated here for the tags such a file holds.  What the reference does with it is pinned by its
  call site: the root compound's children are mesh compounds, each with TAG_Byte_Array "vertices" and "normals" whose bytes are
  little-endian float triplets; every *vertex* (not triangle) pushes a material index 0.
* `save_exr` / `load_exr` — SDK/sutil/sutil.cpp:660-702: float3 / float4 buffers are handed to tinyexr's SaveEXR(…, save_as_fp16 = true),
  which writes a single-part scanline file with HALF channels in alphabetical order (A, B, G, R), ZIP-compressed in blocks of 16
  lines when the picture is at least 16 x 16 and uncompressed otherwise, rows top to bottom in buffer order (no vertical flip, unlike
  the PPM / PNG branches)."""
import gzip
import struct
import zlib

import numpy as np

# ---- NBT ----------------------------------------------------------------------------------------------------------------------------
TAG_END, TAG_BYTE, TAG_SHORT, TAG_INT, TAG_LONG, TAG_FLOAT, TAG_DOUBLE, TAG_BYTE_ARRAY, TAG_STRING, TAG_LIST, TAG_COMPOUND, TAG_INT_ARRAY, \
    TAG_LONG_ARRAY = range(13)


class _Reader:
    def __init__(self, data):
        self.d, self.p = data, 0

    def take(self, n):
        if self.p + n > len(self.d):
            raise ValueError("NBT: unexpected end of data")
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def num(self, fmt):
        return struct.unpack(">" + fmt, self.take(struct.calcsize(fmt)))[0]

    def string(self):
        return self.take(self.num("H")).decode("utf-8", "replace")

    def payload(self, tag):
        if tag == TAG_BYTE: return self.num("b")
        if tag == TAG_SHORT: return self.num("h")
        if tag == TAG_INT: return self.num("i")
        if tag == TAG_LONG: return self.num("q")
        if tag == TAG_FLOAT: return self.num("f")
        if tag == TAG_DOUBLE: return self.num("d")
        if tag == TAG_BYTE_ARRAY:
            n = self.num("i")
            if n < 0:
                raise ValueError("NBT: negative array length")
            return np.frombuffer(self.take(n), np.uint8)
        if tag == TAG_STRING: return self.string()
        if tag == TAG_LIST:
            et, n = self.num("b"), self.num("i")
            return [self.payload(et) for _ in range(max(n, 0))]
        if tag == TAG_COMPOUND:
            out = {}  # insertion-ordered: the reference iterates the meshes in file order
            while True:
                t = self.num("b")
                if t == TAG_END:
                    return out
                name = self.string()
                out[name] = self.payload(t)
        if tag == TAG_INT_ARRAY:
            n = self.num("i")
            return np.frombuffer(self.take(4 * max(n, 0)), ">i4")
        if tag == TAG_LONG_ARRAY:
            n = self.num("i")
            return np.frombuffer(self.take(8 * max(n, 0)), ">i8")
        raise ValueError(f"NBT: unknown tag id {tag}")


def read_nbt(path):
    """Parse an NBT file (raw, gzip- or zlib-wrapped) into (root name, nested dict / list / numpy structure)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    elif data[:1] == b"\x78":
        data = zlib.decompress(data)
    r = _Reader(data)
    tag = r.num("b")
    if tag != TAG_COMPOUND:
        raise ValueError("NBT: the root tag is not a compound")
    name = r.string()
    return name, r.payload(TAG_COMPOUND)


def load_nbt(path):
    """imgui_test's load_nbt (SDK/imgui_test/triangle_gas.cpp:16-76): returns (vertices (V, 3) f32, normals (V, 3) f32, mat_indices (V,) i32) —
    the three arrays TriangleGAS keeps, unindexed (three consecutive vertices per triangle).  The reference pushes one material index per
    *vertex* (all 0); a launch only reads the first V / 3 of them.  Missing file: the reference throws "can't find file"."""
    import os
    if not os.path.exists(path):
        raise RuntimeError("can't find file")
    _, root = read_nbt(path)
    verts, norms = [], []
    for mesh in root.values():
        if not isinstance(mesh, dict):
            raise ValueError("NBT model: the root's children must be mesh compounds")
        vb, nb = mesh["vertices"], mesh["normals"]
        n = vb.size // 12            # nvertices = vertex_data.size() / sizeof(float3), remainder bytes ignored
        if nb.size < 12 * n:
            raise ValueError("NBT model: fewer normals than vertices")
        verts.append(np.frombuffer(vb[:12 * n].tobytes(), "<f4").reshape(n, 3))
        norms.append(np.frombuffer(nb[:12 * n].tobytes(), "<f4").reshape(n, 3))
    v = np.concatenate(verts).astype(np.float32) if verts else np.zeros((0, 3), np.float32)
    nrm = np.concatenate(norms).astype(np.float32) if norms else np.zeros((0, 3), np.float32)
    return v, nrm, np.zeros(v.shape[0], np.int32)


def save_nbt(path, meshes, compress=True, root_name=""):
    """Write a model file load_nbt reads: meshes = {name: (vertices (V, 3), normals (V, 3))}.  For tests and for converting models."""
    def name_bytes(s):
        b = s.encode()
        return struct.pack(">H", len(b)) + b
    out = bytearray(struct.pack(">b", TAG_COMPOUND) + name_bytes(root_name))
    for name, (v, n) in meshes.items():
        out += struct.pack(">b", TAG_COMPOUND) + name_bytes(name)
        for key, arr in (("vertices", v), ("normals", n)):
            raw = np.ascontiguousarray(arr, "<f4").tobytes()
            out += struct.pack(">b", TAG_BYTE_ARRAY) + name_bytes(key) + struct.pack(">i", len(raw)) + raw
        out += struct.pack(">b", TAG_END)
    out += struct.pack(">b", TAG_END)
    with open(path, "wb") as f:
        f.write(gzip.compress(bytes(out)) if compress else bytes(out))


# ---- EXR ----------------------------------------------------------------------------------------------------------------------------
_EXR_MAGIC = 20000630
_NO_COMPRESSION, _ZIP_COMPRESSION = 0, 3


def _attr(name, typ, value):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value


def save_exr(path, image):
    """SaveEXR(data, w, h, 3 | 4, save_as_fp16 = 1, …) as sutil::saveImage calls it (SDK/sutil/sutil.cpp:670-696): `image` is a
    (h, w, 3 | 4) float32 buffer, written in buffer order."""
    a = np.asarray(image, np.float32)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError("sutil::saveImage: Unrecognized image buffer pixel format.")
    h, w, nc = a.shape
    names = ["A", "B", "G", "R"] if nc == 4 else ["B", "G", "R"]
    src = {"R": 0, "G": 1, "B": 2, "A": 3}
    half = a.astype(np.float16)
    comp = _ZIP_COMPRESSION if (w >= 16 and h >= 16) else _NO_COMPRESSION
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", 1, 0, 1, 1) for n in names) + b"\0"
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<ii", _EXR_MAGIC, 2)
    hdr += _attr("channels", "chlist", chlist) + _attr("compression", "compression", bytes([comp]))
    hdr += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box) + _attr("lineOrder", "lineOrder", b"\0")
    hdr += _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0))
    hdr += _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    lines_per_chunk = 16 if comp == _ZIP_COMPRESSION else 1
    chunks = []
    for y0 in range(0, h, lines_per_chunk):
        rows = half[y0:y0 + lines_per_chunk]
        # per scanline: every channel's row in turn, alphabetical channel order
        raw = np.stack([rows[:, :, src[n]] for n in names], axis=1).astype("<f2").tobytes()
        data = raw
        if comp == _ZIP_COMPRESSION:
            b = np.frombuffer(raw, np.uint8)
            t = np.concatenate([b[0::2], b[1::2]]).astype(np.int16)            # even bytes, then odd bytes
            t[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xff                        # delta predictor
            z = zlib.compress(t.astype(np.uint8).tobytes())
            data = z if len(z) < len(raw) else raw
        chunks.append(struct.pack("<ii", y0, len(data)) + data)
    table_pos = len(hdr)
    pos = table_pos + 8 * len(chunks)
    table = b""
    for c in chunks:
        table += struct.pack("<Q", pos)
        pos += len(c)
    with open(path, "wb") as f:
        f.write(hdr + table + b"".join(chunks))


def load_exr(path):
    """Read back a scanline EXR of HALF / FLOAT channels with no or ZIP / ZIPS compression (what save_exr and tinyexr's SaveEXR write):
    returns {channel name: (h, w) float32}.  For round-trip tests and for reading the reference's own outputs."""
    with open(path, "rb") as f:
        d = f.read()
    magic, version = struct.unpack_from("<ii", d, 0)
    if magic != _EXR_MAGIC or (version & 0xff) != 2 or (version & 0x1a00):
        raise ValueError("not a single-part scanline OpenEXR file")
    p, attrs = 8, {}
    while d[p] != 0:
        e = d.index(b"\0", p); name = d[p:e].decode(); p = e + 1
        e = d.index(b"\0", p); p = e + 1
        size = struct.unpack_from("<i", d, p)[0]; p += 4
        attrs[name] = d[p:p + size]; p += size
    p += 1
    chans, q, cl = [], 0, attrs["channels"]
    while cl[q] != 0:
        e = cl.index(b"\0", q)
        ptype, = struct.unpack_from("<i", cl, e + 1)
        chans.append((cl[q:e].decode(), ptype))
        q = e + 1 + 16
    comp = attrs["compression"][0]
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines_per_chunk = {0: 1, 2: 1, 3: 16}.get(comp)
    if lines_per_chunk is None:
        raise ValueError(f"EXR compression {comp} not supported")
    nchunks = (h + lines_per_chunk - 1) // lines_per_chunk
    offsets = struct.unpack_from(f"<{nchunks}Q", d, p)
    bpp = {1: 2, 2: 4}
    line_bytes = sum(bpp[t] for _, t in chans) * w
    out = {n: np.zeros((h, w), np.float32) for n, _ in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", d, off)
        data = d[off + 8:off + 8 + size]
        nl = min(lines_per_chunk, y1 - y + 1)
        if comp != 0 and size < line_bytes * nl:
            t = np.frombuffer(zlib.decompress(data), np.uint8).astype(np.int64)
            t = ((np.cumsum(t - 128) + 128) & 0xff).astype(np.uint8)           # undo the delta predictor (first byte is stored as is)
            half_n = (t.size + 1) // 2
            b = np.empty(t.size, np.uint8)
            b[0::2], b[1::2] = t[:half_n], t[half_n:]
            data = b.tobytes()
        q = 0
        for ly in range(nl):
            for n, t in chans:
                nb = bpp[t] * w
                out[n][y - y0 + ly] = np.frombuffer(data[q:q + nb], "<f2" if t == 1 else "<f4").astype(np.float32)
                q += nb
    return out

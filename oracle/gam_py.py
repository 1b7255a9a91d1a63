"""TEST INFRASTRUCTURE ONLY (oracle side) -- pure-Python reader for vg GAM streams.

Restates, for the checker, the wire format the reference consumes through
`vg::io::for_each<vg::Alignment>` (reference src/readGAM.h:20-68) -- vg / libvgio
are third-party and absent from /root/reference, so the format is stated from the
bytes of the reference's own fixtures:

  stream  := gzip/BGZF( group* )
  group   := varint count, then `count` items, each varint length + bytes;
             item 0 is the type tag "GAM", the rest are serialized vg.Alignment.
  Alignment: 1 sequence, 2 path, 3 name, 4 quality (raw PHRED bytes), 5 mapping_quality,
             16 identity (fixed64 double), 20 read_paired
  Path: 2 mapping*   Mapping: 1 position, 2 edit*, 5 rank
  Position: 1 node_id, 2 offset, 4 is_reverse
  Edit: 1 from_length, 2 to_length, 3 sequence

Nothing under vgan_b200/ may import this module.
"""
import gzip
import struct


def _varint(buf, i):
    shift = 0
    val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if not (b & 0x80):
            return val, i
        shift += 7


def _fields(buf):
    """Yield (field_no, wire_type, value) for one serialized message."""
    i = 0
    n = len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = buf[i:i + 8]
            i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = buf[i:i + ln]
            i += ln
        elif wt == 5:
            v = buf[i:i + 4]
            i += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fno, wt, v


def _edit(buf):
    e = {"from_length": 0, "to_length": 0, "sequence": b""}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            e["from_length"] = v
        elif fno == 2:
            e["to_length"] = v
        elif fno == 3:
            e["sequence"] = bytes(v)
    return e


def _mapping(buf):
    m = {"node_id": 0, "offset": 0, "is_reverse": False, "edits": []}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    m["node_id"] = v2
                elif f2 == 2:
                    m["offset"] = v2
                elif f2 == 4:
                    m["is_reverse"] = bool(v2)
        elif fno == 2:
            m["edits"].append(_edit(v))
    return m


def _alignment(buf):
    a = {"sequence": b"", "quality": b"", "mapping_quality": 0, "identity": 0.0,
         "read_paired": False, "name": b"", "mappings": []}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            a["sequence"] = bytes(v)
        elif fno == 2:
            for f2, w2, v2 in _fields(v):
                if f2 == 2:
                    a["mappings"].append(_mapping(v2))
        elif fno == 3:
            a["name"] = bytes(v)
        elif fno == 4:
            a["quality"] = bytes(v)
        elif fno == 5:
            a["mapping_quality"] = v
        elif fno == 16:
            a["identity"] = struct.unpack("<d", v)[0]
        elif fno == 20:
            a["read_paired"] = bool(v)
    return a


def read_gam_bytes(raw):
    """Decode an (already decompressed) GAM byte string into a list of alignment dicts."""
    out = []
    i = 0
    n = len(raw)
    while i < n:
        count, i = _varint(raw, i)
        for k in range(count):
            ln, i = _varint(raw, i)
            item = raw[i:i + ln]
            i += ln
            if k == 0:
                if item != b"GAM":
                    raise ValueError("unexpected type tag %r" % item)
                continue
            out.append(_alignment(item))
    return out


def read_gam(path):
    with open(path, "rb") as fh:
        head = fh.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as fh:
            raw = fh.read()
    else:
        with open(path, "rb") as fh:
            raw = fh.read()
    return read_gam_bytes(raw)


def read_gfa(path):
    """Return ({node_id: sequence}, [(path_name, [(node_id, is_reverse), ...])]) from a GFA1 file."""
    nodes = {}
    paths = []
    with open(path) as fh:
        for line in fh:
            t = line.rstrip("\n").split("\t")
            if t[0] == "S":
                nodes[int(t[1])] = t[2]
            elif t[0] == "P":
                steps = [(int(s[:-1]), s[-1] == "-") for s in t[2].split(",")]
                paths.append((t[1], steps))
    return nodes, paths

"""Minimal pure-Python Avro container reader/writer for tests (null + deflate codecs).  Independent of the C++
codec in ml-ease_b200/host/avro_io.cpp, so the two check each other."""
import json
import os
import struct
import zlib


class _Buf:
    def __init__(self, b):
        self.b, self.i = b, 0

    def long(self):
        shift, acc = 0, 0
        while True:
            c = self.b[self.i]
            self.i += 1
            acc |= (c & 0x7F) << shift
            if not c & 0x80:
                break
            shift += 7
        return (acc >> 1) ^ -(acc & 1)

    def bytes_(self):
        n = self.long()
        out = self.b[self.i:self.i + n]
        self.i += n
        return out

    def raw(self, n):
        out = self.b[self.i:self.i + n]
        self.i += n
        return out


def _named(schema, table):
    if isinstance(schema, dict):
        if schema.get("type") in ("record", "enum", "fixed"):
            table[schema["name"]] = schema
        for f in schema.get("fields", []):
            _named(f["type"], table)
        if "items" in schema:
            _named(schema["items"], table)
        if isinstance(schema.get("type"), (dict, list)):
            _named(schema["type"], table)
    elif isinstance(schema, list):
        for s in schema:
            _named(s, table)


def _decode(buf, schema, table):
    if isinstance(schema, list):
        return _decode(buf, schema[buf.long()], table)
    if isinstance(schema, dict):
        t = schema["type"]
        if t == "record":
            return {f["name"]: _decode(buf, f["type"], table) for f in schema["fields"]}
        if t == "array":
            out = []
            while True:
                n = buf.long()
                if n == 0:
                    break
                if n < 0:
                    n = -n
                    buf.long()
                for _ in range(n):
                    out.append(_decode(buf, schema["items"], table))
            return out
        return _decode(buf, t, table)
    if schema in table:
        return _decode(buf, table[schema], table)
    if schema == "null":
        return None
    if schema == "string":
        return buf.bytes_().decode()
    if schema in ("int", "long"):
        return buf.long()
    if schema == "float":
        return struct.unpack("<f", buf.raw(4))[0]
    if schema == "double":
        return struct.unpack("<d", buf.raw(8))[0]
    if schema == "boolean":
        return buf.raw(1) != b"\0"
    raise ValueError(schema)


def read_avro(path):
    """-> (schema dict, list of records, number of blocks)."""
    b = _Buf(open(path, "rb").read())
    assert b.raw(4) == b"Obj\x01"
    meta = {}
    while True:
        n = b.long()
        if n == 0:
            break
        for _ in range(abs(n)):
            k = b.bytes_().decode()
            meta[k] = b.bytes_()
    codec = meta.get("avro.codec", b"null")
    schema = json.loads(meta["avro.schema"])
    table = {}
    _named(schema, table)
    sync = b.raw(16)
    recs, nblocks = [], 0
    while b.i < len(b.b):
        cnt = b.long()
        size = b.long()
        payload = b.raw(size)
        if codec == b"deflate":
            payload = zlib.decompress(payload, -15)
        pb = _Buf(payload)
        for _ in range(cnt):
            recs.append(_decode(pb, schema, table))
        assert b.raw(16) == sync
        nblocks += 1
    return schema, recs, nblocks


def read_dir(path):
    recs = []
    files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(path) for f in fs if not f.startswith(("_", ".")))
    for f in files:
        recs += read_avro(f)[1]
    return recs


def _wlong(out, v):
    z = (v << 1) ^ (v >> 63)
    while z & ~0x7F:
        out.append((z & 0x7F) | 0x80)
        z >>= 7
    out.append(z)


def _encode(out, schema, v, table):
    if isinstance(schema, list):
        for i, s in enumerate(schema):
            nm = s if isinstance(s, str) else s.get("type")
            if (v is None) == (nm == "null"):
                _wlong(out, i)
                return _encode(out, s, v, table)
        raise ValueError("no union branch for %r" % (v,))
    if isinstance(schema, dict):
        t = schema["type"]
        if t == "record":
            for f in schema["fields"]:
                _encode(out, f["type"], v[f["name"]], table)
            return
        if t == "array":
            if v:
                _wlong(out, len(v))
                for e in v:
                    _encode(out, schema["items"], e, table)
            _wlong(out, 0)
            return
        return _encode(out, t, v, table)
    if schema in table:
        return _encode(out, table[schema], v, table)
    if schema == "null":
        return
    if schema == "string":
        b = v.encode()
        _wlong(out, len(b))
        out.extend(b)
    elif schema in ("int", "long"):
        _wlong(out, int(v))
    elif schema == "float":
        out.extend(struct.pack("<f", v))
    elif schema == "double":
        out.extend(struct.pack("<d", v))
    elif schema == "boolean":
        out.append(1 if v else 0)
    else:
        raise ValueError(schema)


def write_avro(path, schema, records, codec="null", block=100):
    table = {}
    _named(schema, table)
    out = bytearray(b"Obj\x01")
    _wlong(out, 2)
    for k, v in (("avro.schema", json.dumps(schema).encode()), ("avro.codec", codec.encode())):
        _wlong(out, len(k)); out.extend(k.encode()); _wlong(out, len(v)); out.extend(v)
    _wlong(out, 0)
    sync = bytes(range(16))
    out.extend(sync)
    for s in range(0, len(records), block):
        body = bytearray()
        chunk = records[s:s + block]
        for r in chunk:
            _encode(body, schema, r, table)
        if codec == "deflate":
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = co.compress(bytes(body)) + co.flush()
        _wlong(out, len(chunk)); _wlong(out, len(body)); out.extend(body); out.extend(sync)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    open(path, "wb").write(bytes(out))


# the Pig-generated schema of the reference's examples/sample-data.avro
PIG_SCHEMA = {"type": "record", "name": "TUPLE_0", "fields": [
    {"name": "features", "type": ["null", {"type": "array", "items": ["null", {"type": "record", "name": "TUPLE_1", "fields": [
        {"name": "name", "type": ["null", "string"]}, {"name": "term", "type": ["null", "string"]},
        {"name": "value", "type": ["null", "float"]}]}]}]},
    {"name": "offset", "type": ["null", "int"]}, {"name": "response", "type": ["null", "int"]},
    {"name": "weight", "type": ["null", "int"]}]}


def fixture_records(npz, with_key=None):
    """Raw (unprepared) records of the decoded fixture; with_key(i) -> value of an extra int field 'pkey'."""
    names = [str(n) for n in npz["feature_names"]]
    recs = []
    rp = npz["rowptr"]
    for i in range(len(npz["response"])):
        feats = [{"name": names[c], "term": "", "value": float(v)} for c, v in
                 zip(npz["colidx"][rp[i]:rp[i + 1]], npz["val"][rp[i]:rp[i + 1]])]
        r = {"features": feats, "offset": int(npz["offset"][i]), "response": int(npz["response"][i]), "weight": int(npz["weight"][i])}
        if with_key is not None:
            r["pkey"] = int(with_key(i))
        recs.append(r)
    return recs


def pig_schema_with_key():
    s = json.loads(json.dumps(PIG_SCHEMA))
    s["fields"].append({"name": "pkey", "type": ["null", "int"]})
    return s

"""A dependency-free reader for the ONNX model files RapidDoc ships (`rapid_doc/resources/*.onnx`): protobuf wire format ->
nodes / initializers / graph inputs-outputs.  The `onnx` and `onnxruntime` packages are not available on the target boxes;
only the small subset of the schema these CNN graphs use is decoded (ModelProto.graph, NodeProto, AttributeProto, TensorProto
with raw_data / float_data / int64_data, ValueInfoProto names)."""
import struct

import numpy as np

_DT = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def _fields(b):
    """Yield (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v = bytes(b[i:i + 8]); i += 8
        elif w == 2:
            ln, i = _varint(b, i)
            v = b[i:i + ln]; i += ln
        elif w == 5:
            v = bytes(b[i:i + 4]); i += 4
        else:
            raise ValueError(f"unsupported wire type {w}")
        yield f, w, v


def _packed_varints(v):
    out, i = [], 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(x - (1 << 64) if x >= (1 << 63) else x)
    return out


def _tensor(b):
    dims, dt, raw, name, floats, ints = [], 1, None, "", [], []
    for f, w, v in _fields(b):
        if f == 1:
            dims += _packed_varints(v) if w == 2 else [v]
        elif f == 2:
            dt = v
        elif f == 4:
            floats += list(np.frombuffer(bytes(v), np.float32)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 7:
            ints += _packed_varints(v) if w == 2 else [v - (1 << 64) if v >= (1 << 63) else v]
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
    if raw is not None:
        a = np.frombuffer(raw, _DT[dt]).copy()
    elif floats:
        a = np.array(floats, np.float32)
    else:
        a = np.array(ints, _DT.get(dt, np.int64))
    return name, a.reshape(dims) if dims else (a.reshape(()) if a.size == 1 else a)


def _attr(b):
    name, val, ints, floats, kind = "", None, [], [], None
    for f, w, v in _fields(b):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            val = struct.unpack("<f", v)[0]
        elif f == 3:
            val = v - (1 << 64) if v >= (1 << 63) else v
        elif f == 4:
            val = bytes(v).decode("utf-8", "replace")
        elif f == 5:
            val = _tensor(v)[1]
        elif f == 6:
            val = _graph(v)
        elif f == 7:
            floats += list(np.frombuffer(bytes(v), np.float32)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 8:
            ints += _packed_varints(v) if w == 2 else [v]
        elif f == 20:
            kind = v
    if kind == 7 or (val is None and ints):
        val = ints
    elif kind == 6 or (val is None and floats):
        val = floats
    return name, val


class Node:
    def __init__(self, op, inputs, outputs, attrs, name):
        self.op, self.inputs, self.outputs, self.attrs, self.name = op, inputs, outputs, attrs, name

    def __repr__(self):
        return f"{self.op}({', '.join(self.inputs)}) -> {', '.join(self.outputs)} {self.attrs if self.attrs else ''}"


class Graph:
    def __init__(self, nodes, init, inputs, outputs, meta=None):
        self.nodes, self.init, self.inputs, self.outputs, self.meta = nodes, init, inputs, outputs, meta or {}


def _graph(graph, meta=None):
    nodes, init, inputs, outputs = [], {}, [], []
    for f, w, v in _fields(graph):
        if f == 1:
            ins, outs, attrs, op, name = [], [], {}, "", ""
            for g, _, x in _fields(v):
                if g == 1:
                    ins.append(bytes(x).decode())
                elif g == 2:
                    outs.append(bytes(x).decode())
                elif g == 3:
                    name = bytes(x).decode()
                elif g == 4:
                    op = bytes(x).decode()
                elif g == 5:
                    k, val = _attr(x)
                    attrs[k] = val
            nodes.append(Node(op, ins, outs, attrs, name))
        elif f == 5:
            n, a = _tensor(v)
            init[n] = a
        elif f in (11, 12):
            for g, _, x in _fields(v):
                if g == 1:
                    (inputs if f == 11 else outputs).append(bytes(x).decode())
    inputs = [i for i in inputs if i not in init]
    return Graph(nodes, init, inputs, outputs, meta)


def load(path):
    data = memoryview(open(path, "rb").read())
    graph, meta = None, {}
    for f, w, v in _fields(data):
        if f == 7:
            graph = v
        elif f == 14:                                   # metadata_props: StringStringEntryProto {1: key, 2: value}
            kv = {g: bytes(x).decode("utf-8", "replace") for g, _, x in _fields(v) if g in (1, 2)}
            meta[kv.get(1, "")] = kv.get(2, "")
    assert graph is not None, "no graph in the model file"
    return _graph(graph, meta)

#!/usr/bin/env python3
"""Text .param -> binary .param.bin, the format Net::load_param_bin reads (reference: src/net.cpp:1667-1940, written by the
reference's tools/ncnn2mem.cpp:160-505): int32 magic 7767517, layer count, blob count; per layer int32 typeindex, bottom count,
top count, the bottom and top blob indices (a blob's index is the order in which it first appears as a top), then the parameters
as (id, value) pairs -- a value is a float32 when its text holds '.' or 'e', else an int32; arrays are id = -23300 - id, length,
values -- closed by -233.  Names are not stored.

    python tools/param2bin.py model.param model.param.bin

`type_index` maps an operator name to its typeindex; the default is the reference's registry order
(ncnn_b200/csrc/host/layer_type_table.h, generated from src/CMakeLists.txt)."""
import os
import re
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def stock_type_index():
    text = open(os.path.join(HERE, "..", "ncnn_b200", "csrc", "host", "layer_type_table.h")).read()
    body = text[text.index("layer_type_names[] = {"):]
    names = re.findall(r'"(\w+)"', body[:body.index("};")])
    return {n: i for i, n in enumerate(names)}


def _is_float(tok):
    return "." in tok or "e" in tok.lower()


def _value(tok):
    return struct.pack("<f", float(tok)) if _is_float(tok) else struct.pack("<i", int(tok))


def convert(text, type_index=None):
    type_index = type_index or stock_type_index()
    lines = [l for l in text.splitlines() if l.strip()]
    if int(lines[0]) != 7767517:
        raise ValueError("not a .param file")
    layer_count, blob_count = (int(v) for v in lines[1].split())
    out = [struct.pack("<iii", 7767517, layer_count, blob_count)]
    blob_index = {}
    for line in lines[2:2 + layer_count]:
        tok = line.split()
        ltype, nb, nt = tok[0], int(tok[2]), int(tok[3])
        if ltype not in type_index:
            raise ValueError("no typeindex for layer type " + ltype)
        out.append(struct.pack("<iii", type_index[ltype], nb, nt))
        for b in tok[4:4 + nb]:
            out.append(struct.pack("<i", blob_index[b]))
        for t in tok[4 + nb:4 + nb + nt]:
            blob_index[t] = len(blob_index)
            out.append(struct.pack("<i", blob_index[t]))
        for kv in tok[4 + nb + nt:]:
            k, v = kv.split("=", 1)
            key = int(k)
            vals = v.split(",")
            if key <= -23300:      # old array syntax: -233xx=len,v0,v1,...
                out.append(struct.pack("<ii", key, int(vals[0])))
                out.extend(_value(x) for x in vals[1:1 + int(vals[0])])
            elif len(vals) > 1:    # new array syntax: id=v0,v1,... (typed by the first value, tools/ncnn2mem.cpp:406-480)
                isf = _is_float(vals[0])
                out.append(struct.pack("<ii", -key - 23300, len(vals)))
                out.extend(struct.pack("<f", float(x)) if isf else struct.pack("<i", int(x)) for x in vals)
            else:
                out.append(struct.pack("<i", key) + _value(v))
        out.append(struct.pack("<i", -233))
    if len(blob_index) != blob_count:
        raise ValueError("blob count in the header (%d) != blobs found (%d)" % (blob_count, len(blob_index)))
    return b"".join(out)


if __name__ == "__main__":
    data = convert(open(sys.argv[1]).read())
    open(sys.argv[2], "wb").write(data)
    print("%s: %d bytes" % (sys.argv[2], len(data)))

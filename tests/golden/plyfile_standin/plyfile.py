"""A stand-in for the `plyfile` package (absent in this image), just large enough for the reference's
scene/gaussian_model.py save_ply / load_ply: PlyElement.describe, PlyData(list).write, PlyData.read, element[name], .count.
It writes what plyfile writes for these calls — "format binary_little_endian 1.0", one `element <name> <count>` block per
element with `property <type> <name>` lines in dtype order (type names char/uchar/short/ushort/int/uint/float/double),
`end_header`, then each element's packed little-endian records — and nothing else (no comments, no obj_info).
Test infrastructure: used only by tests/golden/make_golden_ply.py to run the REFERENCE's own writer / reader here."""
import numpy as np

_TYPE = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}
_NP = {v: k for k, v in _TYPE.items()}


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    @property
    def count(self):
        return int(self.data.shape[0])

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements=()):
        self.elements = list(elements)

    def write(self, path):
        head = ["ply", "format binary_little_endian 1.0"]
        for e in self.elements:
            head.append(f"element {e.name} {e.count}")
            for field in e.data.dtype.names:
                dt = e.data.dtype.fields[field][0]
                head.append(f"property {_TYPE[dt.kind + str(dt.itemsize)]} {field}")
        head.append("end_header")
        with open(path, "wb") as f:
            f.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                f.write(np.ascontiguousarray(e.data.astype(e.data.dtype.newbyteorder("<"), copy=False)).tobytes())

    @staticmethod
    def read(path):
        raw = open(path, "rb").read()
        end = raw.index(b"end_header")
        body = raw.index(b"\n", end) + 1
        layout = []
        for ln in raw[:end].decode("ascii").split("\n")[2:]:
            tok = ln.split()
            if tok and tok[0] == "element":
                layout.append((tok[1], int(tok[2]), []))
            elif tok and tok[0] == "property":
                layout[-1][2].append((tok[2], "<" + _NP[tok[1]]))
        out, off = [], body
        for name, count, props in layout:
            dt = np.dtype(props)
            out.append(PlyElement(name, np.frombuffer(raw, dtype=dt, count=count, offset=off)))
            off += dt.itemsize * count
        return PlyData(out)

"""oracle/stubs: plyfile stand-in — binary little-endian PLY writer / reader for structured numpy arrays, enough for
GaussianModel.save_ply / load_ply (scene/gaussian_model.py:619-636, 350-385) and dataset_readers.storePly / fetchPly
(`scene.ply` = one `vertex` element, float32 properties x y z nx ny nz f_dc_* f_rest_* opacity scale_*
rot_*).  Own code; the interface follows the plyfile package the reference imports."""
import numpy as np

_TYPES = {"f4": "float", "f8": "double", "u1": "uchar", "i1": "char", "u2": "ushort", "i2": "short", "u4": "uint", "i4": "int"}
_REV = {v: k for k, v in _TYPES.items()}
_REV.update({"float32": "f4", "float64": "f8", "uint8": "u1", "int8": "i1", "uint16": "u2", "int16": "i2", "uint32": "u4", "int32": "i4"})


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name, **_):
        data = np.asarray(data)
        if data.dtype.names is None:
            raise ValueError("PlyElement.describe needs a structured array")
        return PlyElement(name, data)

    def __getitem__(self, key):
        return self.data[key]

    @property
    def properties(self):
        class _P:
            def __init__(self, n):
                self.name = n

        return [_P(n) for n in self.data.dtype.names]

    @property
    def count(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="<"):
        self.elements = list(elements)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, path):
        with open(path, "wb") as f:
            head = ["ply", "format binary_little_endian 1.0"]
            for e in self.elements:
                head.append(f"element {e.name} {len(e.data)}")
                for n in e.data.dtype.names:
                    dt = e.data.dtype[n]
                    head.append(f"property {_TYPES[dt.str.lstrip('<>=|')]} {n}")
            head.append("end_header")
            f.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                le = e.data.astype(e.data.dtype.newbyteorder("<"), copy=False)
                f.write(np.ascontiguousarray(le).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            fmt = f.readline().decode().split()
            elems, cur = [], None
            while True:
                line = f.readline().decode().strip()
                if line == "end_header":
                    break
                t = line.split()
                if t[0] == "element":
                    cur = [t[1], int(t[2]), []]
                    elems.append(cur)
                elif t[0] == "property":
                    cur[2].append((t[2], _REV[t[1]]))
            out = []
            if fmt[1] == "ascii":
                for name, n, props in elems:
                    arr = np.zeros(n, dtype=[(p, "<" + d) for p, d in props])
                    for i in range(n):
                        vals = f.readline().decode().split()
                        for (p, d), v in zip(props, vals):
                            arr[p][i] = float(v)
                    out.append(PlyElement(name, arr))
            else:
                bo = "<" if "little" in fmt[1] else ">"
                for name, n, props in elems:
                    dt = np.dtype([(p, bo + d) for p, d in props])
                    out.append(PlyElement(name, np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n).copy()))
        return PlyData(out)

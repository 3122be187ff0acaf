"""Torch7 binary serialisation (``torch.save`` / ``torch.load`` default format), SURVEY.md 9.10.

Little-endian; int = 4 B, long = 8 B, double = 8 B.  ``writeObject`` emits a type tag
(0 nil, 1 number, 2 string, 3 table, 4 torch object, 5 boolean); tables and torch objects carry a
1-based reference index shared by both kinds and are written once.  A torch object is
``"V 1"``, its class name, then either the class's own payload (tensors, storages) or -- for
``nn.*`` modules -- one table with all fields.  Field order follows Lua ``pairs()`` and is not
stable, so checkpoints are compared tensor by tensor after loading, never byte by byte.

Python mapping: nil <-> None, number <-> float (ints are written as doubles), string <-> str,
boolean <-> bool, table <-> dict (a list is written as a 1..n table; a loaded table whose keys are
exactly 1..n comes back as a list), tensor <-> numpy array, other torch objects <-> TorchObject.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5

_TENSOR_CLASSES = {
    np.dtype(np.float32): ("torch.FloatTensor", "torch.FloatStorage"),
    np.dtype(np.float64): ("torch.DoubleTensor", "torch.DoubleStorage"),
    np.dtype(np.uint8): ("torch.ByteTensor", "torch.ByteStorage"),
    np.dtype(np.int64): ("torch.LongTensor", "torch.LongStorage"),
    np.dtype(np.int32): ("torch.IntTensor", "torch.IntStorage"),
}
_STORAGE_DTYPES = {v[1]: k for k, v in _TENSOR_CLASSES.items()}
_TENSOR_TO_STORAGE = {v[0]: v[1] for v in _TENSOR_CLASSES.values()}
# CUDA tensors saved by the reference before :float() would appear under these names
_STORAGE_DTYPES["torch.CudaStorage"] = np.dtype(np.float32)
_TENSOR_TO_STORAGE["torch.CudaTensor"] = "torch.CudaStorage"


class Storage(np.ndarray):
    """A torch.*Storage (e.g. the LongStorage held by nn.View.size); a numpy array with a marker type."""

    def __new__(cls, data, dtype=np.int64):
        return np.asarray(data, dtype=dtype).reshape(-1).view(cls)


class TorchObject:
    """A torch class instance that is not a tensor/storage, e.g. ``nn.SpatialConvolution``."""

    def __init__(self, classname, fields=None):
        self.classname = classname
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def __contains__(self, k):
        return k in self.fields

    def __repr__(self):
        return "TorchObject(%s, %s)" % (self.classname, sorted(map(str, self.fields)))


# ------------------------------------------------------------------------------ writer
class _Writer:
    def __init__(self, f):
        self.f = f
        self.index = {}      # id(obj) -> reference index
        self.keep = []       # keep written objects alive so ids stay unique
        self.next = 1

    def int(self, v):
        self.f.write(struct.pack("<i", v))

    def long(self, v):
        self.f.write(struct.pack("<q", v))

    def string(self, s):
        b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
        self.int(len(b))
        self.f.write(b)

    def _ref(self, obj):
        """Write the reference index; return True when the body still has to be written."""
        k = id(obj)
        if k in self.index:
            self.int(self.index[k])
            return False
        self.index[k] = self.next
        self.keep.append(obj)
        self.int(self.next)
        self.next += 1
        return True

    def obj(self, o):
        if o is None:
            self.int(TYPE_NIL)
        elif isinstance(o, bool):
            self.int(TYPE_BOOLEAN)
            self.int(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER)
            self.f.write(struct.pack("<d", float(o)))
        elif isinstance(o, str):
            self.int(TYPE_STRING)
            self.string(o)
        elif isinstance(o, Storage):
            self.int(TYPE_TORCH)
            if self._ref(o):
                self.string("V 1")
                self.string(_TENSOR_CLASSES[o.dtype][1])
                self.long(o.size)
                self.f.write(np.asarray(o).astype(o.dtype.newbyteorder("<"), copy=False).tobytes())
        elif isinstance(o, np.ndarray):
            self.tensor(o)
        elif isinstance(o, TorchObject):
            self.int(TYPE_TORCH)
            if self._ref(o):
                self.string("V 1")
                self.string(o.classname)
                self.obj(o.fields)
        elif isinstance(o, (list, tuple)):
            self.int(TYPE_TABLE)
            if self._ref(o):
                self.int(len(o))
                for i, v in enumerate(o):
                    self.obj(i + 1)
                    self.obj(v)
        elif isinstance(o, dict):
            self.int(TYPE_TABLE)
            if self._ref(o):
                self.int(len(o))
                for k, v in o.items():
                    self.obj(k)
                    self.obj(v)
        else:
            raise TypeError("cannot serialise %r to .t7" % type(o))

    def tensor(self, a):
        if a.dtype not in _TENSOR_CLASSES:
            raise TypeError("no Torch7 tensor type for dtype %s" % a.dtype)
        tcls, scls = _TENSOR_CLASSES[a.dtype]
        self.int(TYPE_TORCH)
        if not self._ref(a):
            return
        self.string("V 1")
        self.string(tcls)
        if a.size == 0:                         # empty tensor: nDim 0, offset 1, nil storage
            self.int(0)
            self.long(1)
            self.int(TYPE_NIL)
            return
        c = np.ascontiguousarray(a)
        self.int(c.ndim)
        for s in c.shape:
            self.long(s)
        for s in c.strides:
            self.long(s // c.itemsize)
        self.long(1)                            # storageOffset, 1-based
        self.int(TYPE_TORCH)                    # the storage object (never shared here: util.save clones tensors)
        self.int(self.next)
        self.next += 1
        self.string("V 1")
        self.string(scls)
        self.long(c.size)
        self.f.write(c.astype(c.dtype.newbyteorder("<"), copy=False).tobytes())


def save(path, obj):
    """torch.save(path, obj)."""
    with open(path, "wb") as f:
        _Writer(f).obj(obj)


# ------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, f):
        self.f = f
        self.objects = {}
        self.stale = {}            # id(dict registered while reading) -> (dict, the list it became): patched by load()

    def _read(self, fmt, n):
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("truncated .t7 file")
        return struct.unpack(fmt, b)[0]

    def int(self):
        return self._read("<i", 4)

    def long(self):
        return self._read("<q", 8)

    def string(self):
        n = self.int()
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("truncated .t7 file")
        return b.decode("utf-8", "replace")

    def obj(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self._read("<d", 8)
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int() == 1
        if t == TYPE_TABLE:
            idx = self.int()
            if idx in self.objects:
                return self.objects[idx]
            d = {}
            self.objects[idx] = d
            n = self.int()
            for _ in range(n):
                k = self.obj()
                d[k] = self.obj()
            # Lua has one table type; here a table whose keys are exactly 1..n is a list, and so is the EMPTY table (the writer emits
            # both [] and {} as a 0-entry table: one fixed mapping keeps a round trip stable).  A back-reference taken while the body
            # was being read (cyclic structures: nngraph gModules) still points at `d`: it is patched after the whole file is read.
            keys = list(d.keys())
            if all(isinstance(k, int) for k in keys) and sorted(keys) == list(range(1, len(keys) + 1)):
                lst = [d[i] for i in range(1, len(keys) + 1)]
                self.objects[idx] = lst
                self.stale[id(d)] = (d, lst)
                return lst
            return d
        if t == TYPE_TORCH:
            idx = self.int()
            if idx in self.objects:
                return self.objects[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version   # pre-"V" files carry the class name first
            if cls in _TENSOR_TO_STORAGE:
                o = self.tensor(cls)
            elif cls in _STORAGE_DTYPES:
                o = self.storage(cls)
            else:
                o = TorchObject(cls)
                self.objects[idx] = o
                fields = self.obj()
                o.fields = fields if isinstance(fields, dict) else ({} if fields == [] else {"_payload": fields})
                return o
            self.objects[idx] = o
            return o
        raise ValueError("unsupported .t7 type tag %d" % t)

    def storage(self, cls):
        n = self.long()
        dt = _STORAGE_DTYPES[cls]
        b = self.f.read(n * dt.itemsize)
        if len(b) != n * dt.itemsize:
            raise EOFError("truncated .t7 storage")
        return np.frombuffer(b, dtype=dt.newbyteorder("<")).astype(dt).view(Storage)

    def tensor(self, cls):
        nd = self.int()
        size = [self.long() for _ in range(nd)]
        stride = [self.long() for _ in range(nd)]
        off = self.long() - 1
        st = self.obj()
        dt = _STORAGE_DTYPES[_TENSOR_TO_STORAGE[cls]]
        if st is None or nd == 0:
            return np.zeros([0], dt)
        return np.lib.stride_tricks.as_strided(np.asarray(st)[off:], shape=size, strides=[s * dt.itemsize for s in stride]).copy()


def _patch_stale(root, stale):
    """Replace every reference to a table dict that turned out to be a list (see _Reader.obj) -- identity and aliasing survive."""
    if not stale:
        return root
    fix = lambda v: stale[id(v)][1] if isinstance(v, dict) and id(v) in stale and stale[id(v)][0] is v else v
    seen, todo = set(), [root]
    while todo:
        o = todo.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, list):
            for i, v in enumerate(o):
                o[i] = fix(v)
            todo.extend(v for v in o if isinstance(v, (list, dict, TorchObject)))
        elif isinstance(o, dict):
            for k in list(o.keys()):
                o[k] = fix(o[k])
            todo.extend(v for v in o.values() if isinstance(v, (list, dict, TorchObject)))
        elif isinstance(o, TorchObject):
            o.fields = fix(o.fields) if isinstance(fix(o.fields), dict) else o.fields
            todo.append(o.fields)
    return fix(root)


def load(path):
    """torch.load(path)."""
    with open(path, "rb") as f:
        r = _Reader(f)
        return _patch_stale(r.obj(), r.stale)

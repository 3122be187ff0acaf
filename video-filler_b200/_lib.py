"""ctypes binding of libcenn.so, generated from the prototypes in ``include/cenn.h``.

This is the Python stand-in for the LuaJIT ``ffi.cdef`` a maintainer would add to the reference
(INTEGRATION.md); the header is the single source of truth for names and argument types.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "cenn.h")
LIB_PATH = os.path.join(_HERE, "csrc", "libcenn.so")


class TrainerConfig(C.Structure):
    """struct cenn_trainer_config (include/cenn.h)."""
    _fields_ = [("variant", C.c_int), ("batchSize", C.c_int), ("fineSize", C.c_int), ("nBottleneck", C.c_int),
                ("nef", C.c_int), ("ngf", C.c_int), ("ndf", C.c_int), ("nc", C.c_int), ("predLen", C.c_int),
                ("overlapPred", C.c_int), ("wtl2", C.c_float), ("weight_nomask", C.c_float), ("wtgdl", C.c_float),
                ("lr", C.c_float), ("beta1", C.c_float), ("precision", C.c_int), ("world_size", C.c_int),
                ("rank", C.c_int), ("dead_dgrad", C.c_int),
                ("noiseGen", C.c_int), ("nz", C.c_int), ("conditionAdv", C.c_int), ("bn_local", C.c_int)]


class InpainterConfig(C.Structure):
    """struct cenn_inpainter_config (include/cenn.h)."""
    _fields_ = [("variant", C.c_int), ("batch", C.c_int), ("fineSize", C.c_int), ("nBottleneck", C.c_int),
                ("nef", C.c_int), ("ngf", C.c_int), ("nc", C.c_int), ("inputLen", C.c_int)]


_TYPES = {
    "int": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "size_t": C.c_size_t, "float": C.c_float,
    "double": C.c_double, "void": None,
}


def _ctype(decl):
    """Map a C parameter/return declaration (without the name) to a ctypes type."""
    d = decl.replace("const", " ").strip()
    d = re.sub(r"\s+", " ", d)
    stars = d.count("*")
    base = d.replace("*", "").strip()
    if base == "char" and stars == 1:
        return C.c_char_p
    if stars == 0:
        return _TYPES[base]
    # every pointer (device or host, opaque handles, out-params) travels as void*
    return C.c_void_p


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every CENN_API prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"CENN_API\s+([\w\s\*]+?)\b(cenn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a, flags=re.S)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


PROTOS = parse_header()
_lib = None


class CennError(RuntimeError):
    pass


def load():
    """dlopen libcenn.so and attach argtypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CennError("libcenn.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C video-filler_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (ret, argtypes, _) in PROTOS.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is not exported
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().cenn_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise CennError("%s failed: %s" % (what or "libcenn call", last_error()))


class Api:
    """Attribute access returns a checked wrapper: ``api.cenn_fill(state, ptr, n, v)`` raises on error."""

    def __init__(self):
        self.lib = load()

    def __getattr__(self, name):
        fn = getattr(self.lib, name)
        ret = PROTOS[name][0]
        if ret is not C.c_int:
            return fn

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise CennError("%s failed: %s" % (name, last_error()))
            return rc
        call.__name__ = name
        setattr(self, name, call)
        return call

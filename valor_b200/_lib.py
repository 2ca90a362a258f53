"""ctypes binding of the C ABI in include/valor_b200.h (built in-tree as csrc/libvalor_b200.so).

The signatures are parsed from the header itself, so the Python side cannot drift from the
ABI.  There is NO fallback: if the shared library is missing the import of any compute path
fails loudly (`ValorLibraryError`).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "valor_b200.h")
LIB_PATH = os.path.join(HERE, "csrc", "libvalor_b200.so")


class ValorLibraryError(RuntimeError):
    pass


class ValorGemmEpilogue(ctypes.Structure):
    _fields_ = [("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("act_aux", ctypes.c_void_p),
                ("preact_out", ctypes.c_void_p), ("ldr", ctypes.c_longlong), ("ld_aux", ctypes.c_longlong),
                ("ld_pre", ctypes.c_longlong), ("res_dtype", ctypes.c_int), ("aux_dtype", ctypes.c_int),
                ("act", ctypes.c_int), ("out_dtype", ctypes.c_int), ("accumulate", ctypes.c_int),
                ("alpha", ctypes.c_float), ("bias_grad", ctypes.c_void_p), ("row_scale", ctypes.c_void_p),
                ("rows_per_group", ctypes.c_int)]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtype...])} for every function the header declares."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"typedef struct .*?\} \w+;", " ", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(valor_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("long long"):
                    argtypes.append(ctypes.c_longlong)
                elif a.startswith("float"):
                    argtypes.append(ctypes.c_float)
                elif a.startswith("int"):
                    argtypes.append(ctypes.c_int)
                else:
                    raise ValorLibraryError(f"cannot parse argument '{a}' of {name}")
        restype = ctypes.c_char_p if ret.startswith("const char") else (ctypes.c_longlong if ret == "long long" else ctypes.c_int)
        decls[name] = (restype, argtypes)
    return decls


_lib = None
_decls = None


def load():
    global _lib, _decls
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ValorLibraryError(
            f"{LIB_PATH} not found: build the sm_100a kernels first (python -c 'import __graft_entry__ as g; "
            "g.build()' or valor_b200/csrc/build.sh). There is no CPU or library fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _decls = parse_header()
    for name, (ret, argtypes) in _decls.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ValorLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


def declared_symbols():
    return sorted(parse_header().keys())


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError with valor_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.valor_last_error().decode()}")

"""ctypes binding of csrc/libtmae_sm100.so.  Argument types are read from include/tmae_sm100.h so the
Python side can never drift from the C ABI.  There is no fallback: a missing library raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtmae_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tmae_sm100.h")

_DECL = re.compile(r"TMAE_API\s+([\w\s\*]+?)\s*\b(tmae_\w+)\s*\(([^;]*?)\)\s*;", re.S)


def _ctype(t):
    t = t.strip()
    if "*" in t:
        return ctypes.c_char_p if t.replace(" ", "") == "constchar*" else ctypes.c_void_p
    t = t.replace("const", "").split()
    base = t[0]
    return {"int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "int": ctypes.c_int, "size_t": ctypes.c_size_t,
            "float": ctypes.c_float, "uint8_t": ctypes.c_uint8, "void": None}[base]


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every TMAE_API declaration."""
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    out = {}
    for ret, name, args in _DECL.findall(src):
        args = args.strip()
        if args in ("", "void"):
            at = []
        else:
            at = []
            for a in args.split(","):
                a = a.strip()
                ty = a if "*" in a else " ".join(a.split()[:-1])  # drop the parameter name
                at.append(_ctype(ty))
        out[name] = (_ctype(ret), at)
    return out


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python t-mae_b200/csrc/build.py` "
                "(tmae_b200 has no CPU or PyTorch fallback)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.decls = parse_header()
        for name, (res, args) in self.decls.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args
            if res is ctypes.c_int and name not in ("tmae_version",):
                setattr(self, name[5:], self._checked(fn, name))
            else:
                setattr(self, name[5:], fn)
        for key, val in os.environ.items():   # TMAE_OPT_<NAME>=<int> -> tmae_set_option (A/B measurements)
            if key.startswith("TMAE_OPT_"):
                self.set_option(key[9:].lower().encode(), int(val))
        if "TMAE_BF16_ATTN_IMPL" in os.environ:
            self.bf16_set_attention_impl(int(os.environ["TMAE_BF16_ATTN_IMPL"]))

    def _checked(self, fn, name):
        err = self.cdll.tmae_last_error_string

        def call(*a):
            r = fn(*a)
            if r != 0:
                raise RuntimeError(f"{name} failed ({r}): {err().decode()}")
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib

"""ctypes binding of the C-ABI library ``libcellbender_b200.so`` (declared in ``include/cellbender_b200.h``).

The prototypes are parsed from the header itself, so the Python side cannot drift from the declared ABI.
There is NO fallback: if the library is missing this module raises, and every op in the package fails loudly.
"""
import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "cellbender_b200.h")
LIB_PATH = os.path.join(HERE, "libcellbender_b200.so")

_SCALARS = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


from ._cabi_counts import LAUNCHES as _LAUNCHES


class CabiError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(cb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if ret == "const char*" else _SCALARS[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise CabiError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m cellbender_b200.build` or `python -c 'import __graft_entry__ as g; g.build()'`."
            )
        self._dll = ctypes.CDLL(LIB_PATH)
        self.launch_count = 0  # kernels launched through this library (bench.py reports it as gpu_launches)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:
                raise CabiError(f"symbol {name} declared in include/cellbender_b200.h is not exported by {LIB_PATH}") from e
            fn.restype = restype
            fn.argtypes = argtypes

    def last_error(self) -> str:
        return self._dll.cb_last_error().decode()

    def call(self, name: str, *args):
        """Call an int-status entry point; raise CabiError with the library's message on failure."""
        self.launch_count += _LAUNCHES.get(name, 1)
        rc = getattr(self._dll, name)(*args)
        if rc != 0:
            raise CabiError(f"{name} failed (status {rc}): {self.last_error()}")

    def raw(self, name: str):
        return getattr(self._dll, name)


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream

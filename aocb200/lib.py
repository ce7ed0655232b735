"""ctypes binding of libaocb200.so -- the C ABI declared in include/aocb200.h.

The prototypes are parsed from the header, so the header stays the single source of truth.  There is no
fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "aocb200.h")
_TAG = os.environ.get("AOCB200_LIB_TAG", "")          # tooling builds only (see build.py); the product library has no tag
LIB_PATH = os.path.join(_HERE, "libaocb200%s.so" % ("_" + _TAG if _TAG else ""))

_CT = {
    "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t, "long long": ctypes.c_longlong,
    "cudaStream_t": ctypes.c_void_p, "unsigned int": ctypes.c_uint,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argtype, argname), ...])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"(?:^|;|\{)\s*((?:const\s+)?[A-Za-z_][\w ]*?[\s\*]+)(aoc_\w+)\s*\(([^)]*)\)\s*(?=;)", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                alist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, alist)
    return protos


def _ctype(t):
    if "*" in t:
        return ctypes.c_char_p if t.replace(" ", "") == "constchar*" else ctypes.c_void_p
    t = t.replace("const ", "").strip()
    return _CT[t]


class AocError(RuntimeError):
    pass


# CUDA kernels launched by one call of each entry point (everything else launches exactly one); used for the
# `gpu_launches` figure of bench.py.
KERNELS_PER_CALL = {
    "aoc_channel_stats_f32": 2, "aoc_affine_stats_nc_f32": 2, "aoc_bank_index_build": 4, "aoc_global_match_simt_f32": 2, "aoc_global_match_tc": 10,
    "aoc_head_pool_f32": 2, "aoc_dyn_logits_f32": 2, "aoc_gemm_tf32x3_test": 3,
    "aoc_global_match_tc_sharded": 9, "aoc_peer_alloc": 0, "aoc_peer_free": 0, "aoc_peer_export": 0, "aoc_peer_open": 0,
    "aoc_peer_close": 0, "aoc_match_shard_range": 0, "aoc_version": 0, "aoc_check_device": 0, "aoc_last_error_string": 0, "aoc_set_option": 0, "aoc_conv_tiles_per_image": 0, "aoc_conv_trace": 0,
}


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise AocError("libaocb200.so is not built (run `python -m aocb200.build`); there is no fallback path")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0          # kernels launched through this binding (see KERNELS_PER_CALL)
        self.calls = 0
        self.profile = None        # {entry point name: [(start_event, end_event, args), ...]} when profiling
        # calls made while a CUDA graph is being captured are bracketed by EXTERNAL events (event-record nodes of the
        # graph, re-recorded by every replay): {entry point name: [(start, end, args, capture tag), ...]}; the engine
        # sets `capture_tag` around a capture and adds the tag of every replayed graph to `replayed`
        self.profile_graph = {}
        self.capture_tag = None
        self.replayed = set()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)
            fn.restype = _ctype(ret) if ret != "int" else ctypes.c_int
            fn.argtypes = [_ctype(t) for t, _ in args]
            if ret == "int" and name not in ("aoc_version", "aoc_conv_tiles_per_image"):
                setattr(self, name[4:], self._checked(fn, name))
            else:
                setattr(self, name[4:], fn)

    def _checked(self, fn, name):
        per_call = KERNELS_PER_CALL.get(name, 1)
        is_query = name.endswith("_bytes")

        def call(*a):
            prof = self.profile
            if prof is not None and name in prof:
                import torch
                cap = torch.cuda.is_current_stream_capturing()
                e0, e1 = (torch.cuda.Event(enable_timing=True, external=cap) for _ in range(2))
                e0.record()
                rc = fn(*a)
                e1.record()
                if cap:
                    self.profile_graph.setdefault(name, []).append((e0, e1, a, self.capture_tag))
                else:
                    prof[name].append((e0, e1, a))
            else:
                rc = fn(*a)
            if rc != 0:
                raise AocError("%s failed (%d): %s" % (name, rc, self.cdll.aoc_last_error_string().decode()))
            if not is_query:
                self.calls += 1
                self.launches += per_call
            return rc
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib

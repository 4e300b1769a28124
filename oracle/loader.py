"""TEST INFRASTRUCTURE -- ctypes loaders for

  * oracle/_ref/libcvtt_ref.so     : the UNMODIFIED reference (built by oracle/Makefile from /root/reference), and
  * oracle/_build/libcvtt_oracle.so: this repo's plain-C restatement of the BC7 path (oracle/cvtt_oracle.c).

Nothing under convectionkernels_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libcvtt_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libcvtt_oracle.so")

# format ids shared with include/cvtt_b200.h (cvttb200_format)
FMT = dict(BC1=1, BC2=2, BC3=3, BC4U=4, BC4S=5, BC5U=6, BC5S=7, BC6HU=8, BC6HS=9, BC7=10,
           ETC1=11, ETC2=12, ETC2_RGBA=13, ETC2_PUNCHTHROUGH=14, ETC2_ALPHA=15, EAC_R11U=16, EAC_R11S=17)
IN_BYTES = {k: 64 for k in FMT}
IN_BYTES.update(BC6HU=128, BC6HS=128, EAC_R11U=32, EAC_R11S=32)
OUT_BYTES = {k: 16 for k in FMT}
OUT_BYTES.update(BC1=8, BC4U=8, BC4S=8, ETC1=8, ETC2=8, ETC2_PUNCHTHROUGH=8, ETC2_ALPHA=8, EAC_R11U=8, EAC_R11S=8)


def build(which=("oracle", "ref")):
    """(Re)build the oracle libraries.  'ref' is a no-op where /root/reference does not exist."""
    for target in which:
        subprocess.check_call(["make", "-s", "-C", HERE, target])


def _as_u8(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


class Reference:
    """The unmodified reference through oracle/ref_shim.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libcvtt_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = ctypes.CDLL(REF_SO)
        for f in ("cvttref_sizeof_options", "cvttref_sizeof_plan", "cvttref_sizeof_finetune"):
            getattr(L, f).restype = ctypes.c_size_t
        L.cvttref_encode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.cvttref_encode_alloc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.cvttref_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.cvttref_rcp.restype = ctypes.c_float
        L.cvttref_rcp.argtypes = [ctypes.c_float]
        L.cvttref_hardware_threads.restype = ctypes.c_uint
        assert L.cvttref_sizeof_options() == 44 and L.cvttref_sizeof_plan() == 808

    def default_options(self):
        buf = np.zeros(44, dtype=np.uint8)
        self.lib.cvttref_default_options(buf.ctypes.data_as(ctypes.c_void_p))
        return buf

    def default_plan(self):
        buf = np.zeros(808, dtype=np.uint8)
        self.lib.cvttref_default_plan(buf.ctypes.data_as(ctypes.c_void_p))
        return buf

    def plan_from_quality(self, q):
        buf = np.zeros(808, dtype=np.uint8)
        self.lib.cvttref_plan_from_quality(buf.ctypes.data_as(ctypes.c_void_p), int(q))
        return buf

    def plan_from_finetune(self, ft):
        ft = np.ascontiguousarray(ft, dtype=np.uint8)
        assert ft.size == self.lib.cvttref_sizeof_finetune()
        buf = np.zeros(808, dtype=np.uint8)
        self.lib.cvttref_plan_from_finetune(buf.ctypes.data_as(ctypes.c_void_p), ft.ctypes.data_as(ctypes.c_void_p))
        return buf

    def encode(self, fmt, blocks, options, plan=None, threads=1, etc2_alloc_options=None):
        """etc2_alloc_options: the Options AllocETC2Data is called with (default: the same as `options`)"""
        src = _as_u8(blocks)
        n = src.size // IN_BYTES[fmt]
        assert n * IN_BYTES[fmt] == src.size and n % 8 == 0
        out = np.zeros((n, OUT_BYTES[fmt]), dtype=np.uint8)
        options = np.ascontiguousarray(options, dtype=np.uint8)
        pp = None if plan is None else np.ascontiguousarray(plan, dtype=np.uint8).ctypes.data_as(ctypes.c_void_p)
        alloc = options if etc2_alloc_options is None else np.ascontiguousarray(etc2_alloc_options, dtype=np.uint8)
        rc = self.lib.cvttref_encode_alloc(FMT[fmt], src.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p),
                                           options.ctypes.data_as(ctypes.c_void_p), pp, alloc.ctypes.data_as(ctypes.c_void_p), int(threads))
        if rc != 0:
            raise ValueError("cvttref_encode rc=%d" % rc)
        return out

    def decode_bc7(self, bc):
        bc = _as_u8(bc)
        n = bc.size // 16
        out = np.zeros((n, 16, 4), dtype=np.uint8)
        rc = self.lib.cvttref_decode(FMT["BC7"], bc.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        return out

    def decode(self, fmt, bc):
        """cvtt::Kernels::DecodeBC7 / DecodeBC6HU / DecodeBC6HS on n encoded blocks (n % 8 == 0)"""
        bc = _as_u8(bc)
        n = bc.size // 16
        assert n % 8 == 0
        out = np.zeros((n, 16, 4), dtype=np.uint8 if fmt == "BC7" else np.int16)
        rc = self.lib.cvttref_decode(FMT[fmt], bc.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise ValueError("cvttref_decode rc=%d" % rc)
        return out

    def rcp(self, v):
        return float(self.lib.cvttref_rcp(ctypes.c_float(v)))

    def rcp_table(self):
        """_mm_rcp_ps(0..16) of this host as float32[17] (entry 0 is inf)."""
        with np.errstate(all="ignore"):
            return np.array([self.rcp(float(n)) for n in range(17)], dtype=np.float32)

    def hardware_threads(self):
        return int(self.lib.cvttref_hardware_threads())


class Oracle:
    """The plain-C restatement (BC7 path)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(("oracle",))
        L = self.lib = ctypes.CDLL(ORACLE_SO)
        L.cvtt_oracle_sizeof_plan.restype = ctypes.c_size_t
        L.cvtt_oracle_sizeof_options.restype = ctypes.c_size_t
        L.cvtt_oracle_encode_bc7.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cvtt_oracle_set_rcp_table.argtypes = [ctypes.c_void_p]
        assert L.cvtt_oracle_sizeof_plan() == 808 and L.cvtt_oracle_sizeof_options() == 44

    def set_rcp_table(self, table):
        """17 floats standing in for _mm_rcp_ps(0..16) (golden fixtures recorded on another CPU), or None."""
        if table is None:
            self.lib.cvtt_oracle_set_rcp_table(None)
        else:
            t = np.ascontiguousarray(table, dtype=np.float32)
            assert t.size == 17
            self.lib.cvtt_oracle_set_rcp_table(t.ctypes.data)

    def plan_from_quality(self, q):
        buf = np.zeros(808, dtype=np.uint8)
        self.lib.cvtt_oracle_plan_from_quality(buf.ctypes.data_as(ctypes.c_void_p), int(q))
        return buf

    def plan_from_finetune(self, ft):
        ft = np.ascontiguousarray(ft, dtype=np.uint8)
        assert ft.size == 285
        buf = np.zeros(808, dtype=np.uint8)
        self.lib.cvtt_oracle_plan_from_finetune(buf.ctypes.data_as(ctypes.c_void_p), ft.ctypes.data_as(ctypes.c_void_p))
        return buf

    def encode_bc7(self, blocks, options, plan):
        src = _as_u8(blocks)
        n = src.size // 64
        assert n % 8 == 0
        out = np.zeros((n, 16), dtype=np.uint8)
        options = np.ascontiguousarray(options, dtype=np.uint8)
        plan = np.ascontiguousarray(plan, dtype=np.uint8)
        rc = self.lib.cvtt_oracle_encode_bc7(src.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p),
                                             options.ctypes.data_as(ctypes.c_void_p), plan.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise ValueError("cvtt_oracle_encode_bc7 rc=%d" % rc)
        return out

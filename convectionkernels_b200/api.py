"""Host-side mirror of the reference's public interface for the encode hot path (reference ConvectionKernels.h:236-277),
over the C ABI of libcvtt_b200.so (include/cvtt_b200.h).

Names follow the reference: Options, BC7EncodingPlan, BC7FineTuningParams, Flags, ConfigureBC7EncodingPlanFromQuality,
EncodeBC7 ... -- except that an Encode* call takes any multiple of NumParallelBlocks (8) blocks instead of exactly 8;
blocks [8k, 8k+8) are the k-th reference call.

Inputs / outputs may be numpy arrays (host memory) or torch CUDA tensors (device memory, work is enqueued on
torch's current stream).  There is no CPU fallback: if the CUDA library or a B200 is missing the call raises.
"""
import ctypes
import os

import numpy as np

from . import build as _build

NumParallelBlocks = 8


class Flags:
    """cvtt::Flags, reference ConvectionKernels.h:33-69"""
    BC7_FastIndexing = 0x008
    BC7_TrySingleColor = 0x010
    BC7_RespectPunchThrough = 0x020
    BC6H_FastIndexing = 0x040
    S3TC_Exhaustive = 0x080
    S3TC_Paranoid = 0x100
    Uniform = 0x200
    ETC_UseFakeBT709 = 0x400
    ETC_FakeBT709Accurate = 0x800
    Fastest = BC6H_FastIndexing | BC7_FastIndexing | S3TC_Paranoid
    Faster = BC6H_FastIndexing | BC7_FastIndexing | S3TC_Paranoid
    Fast = BC7_FastIndexing | S3TC_Paranoid
    Default = BC7_FastIndexing | S3TC_Paranoid
    Better = S3TC_Paranoid | S3TC_Exhaustive
    Ultra = BC7_TrySingleColor | S3TC_Paranoid | S3TC_Exhaustive | ETC_FakeBT709Accurate


class Options(ctypes.Structure):
    """cvtt::Options, reference ConvectionKernels.h:73-103"""
    _fields_ = [("flags", ctypes.c_uint32), ("threshold", ctypes.c_float),
                ("redWeight", ctypes.c_float), ("greenWeight", ctypes.c_float), ("blueWeight", ctypes.c_float), ("alphaWeight", ctypes.c_float),
                ("refineRoundsBC7", ctypes.c_int), ("refineRoundsBC6H", ctypes.c_int), ("refineRoundsIIC", ctypes.c_int),
                ("refineRoundsS3TC", ctypes.c_int), ("seedPoints", ctypes.c_int)]

    def __init__(self, **kw):
        super().__init__()
        _lib().cvttb200_options_default(ctypes.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


class BC7FineTuningParams(ctypes.Structure):
    """cvtt::BC7FineTuningParams, reference ConvectionKernels.h:105-140"""
    _fields_ = [("mode0SP", ctypes.c_uint8 * 16), ("mode1SP", ctypes.c_uint8 * 64), ("mode2SP", ctypes.c_uint8 * 64), ("mode3SP", ctypes.c_uint8 * 64),
                ("mode4SP", (ctypes.c_uint8 * 2) * 4), ("mode5SP", ctypes.c_uint8 * 4), ("mode6SP", ctypes.c_uint8), ("mode7SP", ctypes.c_uint8 * 64)]

    def __init__(self):
        super().__init__()
        _lib().cvttb200_bc7_fine_tuning_default(ctypes.byref(self))


class BC7EncodingPlan(ctypes.Structure):
    """cvtt::BC7EncodingPlan, reference ConvectionKernels.h:142-199"""
    kNumRGBAShapes = 129
    kNumRGBShapes = 243
    _fields_ = [("mode1PartitionEnabled", ctypes.c_uint64), ("mode2PartitionEnabled", ctypes.c_uint64), ("mode3PartitionEnabled", ctypes.c_uint64),
                ("mode0PartitionEnabled", ctypes.c_uint16),
                ("mode7RGBAPartitionEnabled", ctypes.c_uint64), ("mode7RGBPartitionEnabled", ctypes.c_uint64),
                ("mode4SP", (ctypes.c_uint8 * 2) * 4), ("mode5SP", ctypes.c_uint8 * 4), ("mode6Enabled", ctypes.c_uint8),
                ("seedPointsForShapeRGB", ctypes.c_uint8 * 243), ("seedPointsForShapeRGBA", ctypes.c_uint8 * 129),
                ("rgbaShapeList", ctypes.c_uint8 * 129), ("rgbaNumShapesToEvaluate", ctypes.c_uint8),
                ("rgbShapeList", ctypes.c_uint8 * 243), ("rgbNumShapesToEvaluate", ctypes.c_uint8)]

    def __init__(self):
        super().__init__()
        _lib().cvttb200_bc7_plan_default(ctypes.byref(self))

    def tobytes(self):
        return bytes(memoryview(self))


FORMATS = dict(BC1=1, BC2=2, BC3=3, BC4U=4, BC4S=5, BC5U=6, BC5S=7, BC6HU=8, BC6HS=9, BC7=10,
               ETC1=11, ETC2=12, ETC2_RGBA=13, ETC2_PUNCHTHROUGH=14, ETC2_ALPHA=15, EAC_R11U=16, EAC_R11S=17)

_LIB = None


class CvttError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("cvttb200 status %d: %s" % (status, message))
        self.status = status


def library_path():
    return _build.LIB


def _lib():
    """Loads libcvtt_b200.so (building it with nvcc first if it is missing or stale).  Raises if that fails."""
    global _LIB
    if _LIB is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        L.cvttb200_last_error.restype = ctypes.c_char_p
        L.cvttb200_launch_count.restype = ctypes.c_uint64
        L.cvttb200_input_block_bytes.restype = ctypes.c_size_t
        L.cvttb200_output_block_bytes.restype = ctypes.c_size_t
        L.cvttb200_encode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cvttb200_encode_ex.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cvttb200_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.cvttb200_encode_multi.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.cvttb200_set_rcp_table.argtypes = [ctypes.c_void_p]
        L.cvttb200_get_rcp_table.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _check(status):
    if status != 0:
        raise CvttError(status, _lib().cvttb200_last_error().decode("utf-8", "replace"))


def init(device=0):
    _check(_lib().cvttb200_init(int(device)))


def launch_count():
    return int(_lib().cvttb200_launch_count())


def selftest(samples=1 << 28, seed=1):
    """Runs the device self-test (two-lane division vs IEEE division); returns the number of mismatching quotients."""
    bad = ctypes.c_uint64(0)
    L = _lib()
    L.cvttb200_selftest.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    _check(L.cvttb200_selftest(int(samples), int(seed), ctypes.byref(bad)))
    return int(bad.value)


def set_rcp_table(table):
    """table: 17 floats (rcp of 0..16) or None to restore this host's _mm_rcp_ps values."""
    if table is None:
        _check(_lib().cvttb200_set_rcp_table(None))
    else:
        t = np.ascontiguousarray(table, dtype=np.float32)
        assert t.size == 17
        _check(_lib().cvttb200_set_rcp_table(t.ctypes.data))


def get_rcp_table():
    t = np.zeros(17, dtype=np.float32)
    _check(_lib().cvttb200_get_rcp_table(t.ctypes.data))
    return t


def ConfigureBC7EncodingPlanFromQuality(encodingPlan, quality):
    """reference ConvectionKernels.h:262"""
    _lib().cvttb200_bc7_plan_from_quality(ctypes.byref(encodingPlan), int(quality))


def ConfigureBC7EncodingPlanFromFineTuningParams(encodingPlan, params):
    """reference ConvectionKernels.h:265"""
    return bool(_lib().cvttb200_bc7_plan_from_fine_tuning(ctypes.byref(encodingPlan), ctypes.byref(params)))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _as_struct_ptr(obj, ctype):
    if obj is None:
        return None
    if isinstance(obj, ctype):
        return ctypes.byref(obj)
    buf = np.ascontiguousarray(obj, dtype=np.uint8)      # raw bytes (e.g. produced by the oracle loader)
    assert buf.size == ctypes.sizeof(ctype)
    return ctypes.cast(buf.ctypes.data, ctypes.c_void_p), buf


def _check_out(out, like_torch, need_bytes, device=None, dtype="uint8"):
    """A caller-supplied output buffer must be contiguous, of the expected element type, with room for need_bytes, and of the
    same kind (and device) as the input: the library writes through its raw pointer."""
    if _is_torch(out):
        import torch
        if out.dtype != getattr(torch, dtype) or not out.is_contiguous() or out.numel() * out.element_size() < need_bytes:
            raise ValueError("out must be a contiguous %s tensor of at least %d bytes" % (dtype, need_bytes))
        if like_torch and device is not None and out.device != device:
            raise ValueError("out is on %s, the input on %s" % (out.device, device))
        if not like_torch and out.is_cuda:
            raise ValueError("host input needs a host `out` (numpy array or CPU tensor)")
    else:
        if like_torch and device is not None and device.type == "cuda":
            raise ValueError("a CUDA input needs a CUDA tensor as `out`")
        if not isinstance(out, np.ndarray) or out.dtype != np.dtype(dtype) or not out.flags["C_CONTIGUOUS"] or out.nbytes < need_bytes:
            raise ValueError("out must be a C-contiguous %s array of at least %d bytes" % (dtype, need_bytes))
        if not out.flags["WRITEABLE"]:
            raise ValueError("out is read-only")


def encode(fmt, pBlocks, options, encodingPlan=None, out=None, etc2AllocOptions=None):
    """Encodes pBlocks (numpy array or torch CUDA tensor holding n PixelBlocks, n % 8 == 0) to `fmt`.

    etc2AllocOptions: the Options the caller's AllocETC2Data received (ETC2 colour formats only; None = `options`).
    Returns a (n, outBytes) uint8 array of the same kind as the input (or `out` if given)."""
    L = _lib()
    f = FORMATS[fmt] if isinstance(fmt, str) else int(fmt)
    inb, outb = L.cvttb200_input_block_bytes(f), L.cvttb200_output_block_bytes(f)
    if not inb:
        raise ValueError("unknown format %r" % (fmt,))
    keep = []

    def struct_arg(obj, ctype):
        r = _as_struct_ptr(obj, ctype)
        if isinstance(r, tuple):
            keep.append(r[1])
            return r[0]
        return r

    if _is_torch(pBlocks):
        import torch
        src = pBlocks.contiguous()
        nbytes = src.numel() * src.element_size()
        n = nbytes // inb
        if n * inb != nbytes:
            raise ValueError("input size is not a whole number of blocks")
        if out is None:
            out = torch.empty((n, outb), dtype=torch.uint8, device=src.device)
        else:
            _check_out(out, True, n * outb, src.device)
        stream = torch.cuda.current_stream(src.device).cuda_stream if src.is_cuda else None
        with torch.cuda.device(src.device if src.is_cuda else torch.cuda.current_device()):
            st = L.cvttb200_encode_ex(f, src.data_ptr(), n, out.data_ptr(), struct_arg(options, Options), struct_arg(encodingPlan, BC7EncodingPlan),
                                      struct_arg(etc2AllocOptions, Options), stream)
    else:
        src = np.ascontiguousarray(pBlocks)
        nbytes = src.size * src.itemsize
        n = nbytes // inb
        if n * inb != nbytes:
            raise ValueError("input size is not a whole number of blocks")
        if out is None:
            out = np.empty((n, outb), dtype=np.uint8)
        else:
            _check_out(out, False, n * outb)
        optr = out.data_ptr() if _is_torch(out) else out.ctypes.data
        st = L.cvttb200_encode_ex(f, src.ctypes.data, n, optr, struct_arg(options, Options), struct_arg(encodingPlan, BC7EncodingPlan),
                                  struct_arg(etc2AllocOptions, Options), None)
    _check(st)
    return out


def encode_multi(fmt, pBlocks, options, encodingPlan=None, devices=None, out=None):
    """cvttb200_encode_multi: encodes host blocks (numpy, n % 8 == 0) on several GPUs of this process, sharded by whole
    8-block groups; `devices` is a list of device indices (None = every visible device).  Returns a (n, outBytes) uint8 array
    byte-identical to encode() on one device."""
    L = _lib()
    f = FORMATS[fmt] if isinstance(fmt, str) else int(fmt)
    inb, outb = L.cvttb200_input_block_bytes(f), L.cvttb200_output_block_bytes(f)
    keep = []

    def struct_arg(obj, ctype):
        r = _as_struct_ptr(obj, ctype)
        if isinstance(r, tuple):
            keep.append(r[1])
            return r[0]
        return r

    src = np.ascontiguousarray(pBlocks)
    nbytes = src.size * src.itemsize
    n = nbytes // inb
    if n * inb != nbytes:
        raise ValueError("input size is not a whole number of blocks")
    if out is None:
        out = np.empty((n, outb), dtype=np.uint8)
    if devices is None:
        dev_arr, n_dev = None, 0
    else:
        dev_arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        n_dev = len(devices)
    _check(L.cvttb200_encode_multi(f, src.ctypes.data, n, out.ctypes.data, struct_arg(options, Options), struct_arg(encodingPlan, BC7EncodingPlan),
                                   ctypes.cast(dev_arr, ctypes.c_void_p) if dev_arr is not None else None, n_dev))
    return out


def EncodeBC7(pBlocks, options, encodingPlan, out=None):
    """cvtt::Kernels::EncodeBC7, reference ConvectionKernels.h:252 / ConvectionKernels_API.cpp:41-54"""
    return encode("BC7", pBlocks, options, encodingPlan, out)


def EncodeBC6HU(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC6HU, reference ConvectionKernels.h:250 / ConvectionKernels_API.cpp:56-69.  pBlocks: PixelBlockF16
    (int16 half bit patterns, [n][16][4], alpha ignored)."""
    return encode("BC6HU", pBlocks, options, None, out)


def EncodeBC6HS(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC6HS, reference ConvectionKernels.h:251 / ConvectionKernels_API.cpp:71-84"""
    return encode("BC6HS", pBlocks, options, None, out)


def EncodeETC1(pBlocks, options, compressionData=None, out=None):
    """cvtt::Kernels::EncodeETC1, reference ConvectionKernels.h:253 / ConvectionKernels_API.cpp:200-213.  The reference's
    ETC1CompressionData is scratch memory; the library allocates its device equivalent per call (stream-ordered), so the
    argument is accepted for signature compatibility and ignored."""
    return encode("ETC1", pBlocks, options, None, out)


class ETC2CompressionData:
    """What cvtt::Kernels::AllocETC2Data returns (ConvectionKernels.h:268).  The reference's object is CPU scratch plus the chroma
    side axes derived from the allocation-time Options (ConvectionKernels_ETC.cpp:3117-3145); the device scratch is per call
    here, so only the options are kept."""

    def __init__(self, options):
        self.options = Options()
        ctypes.memmove(ctypes.byref(self.options), ctypes.byref(options), ctypes.sizeof(Options))


def AllocETC2Data(options):
    """cvtt::Kernels::AllocETC2Data (the allocator arguments of the reference have no meaning here)"""
    return ETC2CompressionData(options)


def ReleaseETC2Data(compressionData):
    """cvtt::Kernels::ReleaseETC2Data"""
    return None


def _alloc_options(compressionData):
    return None if compressionData is None else compressionData.options


def EncodeETC2(pBlocks, options, compressionData=None, out=None):
    """cvtt::Kernels::EncodeETC2, reference ConvectionKernels.h:254 / ConvectionKernels_API.cpp:215-228.  compressionData: what
    AllocETC2Data returned (None = allocated with the same options)"""
    return encode("ETC2", pBlocks, options, None, out, _alloc_options(compressionData))


def EncodeETC2RGBA(pBlocks, options, compressionData=None, out=None):
    """cvtt::Kernels::EncodeETC2RGBA, reference ConvectionKernels.h:255 / ConvectionKernels_API.cpp:270-286: EAC alpha in bytes
    0-7, ETC2 colour in bytes 8-15"""
    return encode("ETC2_RGBA", pBlocks, options, None, out, _alloc_options(compressionData))


def EncodeETC2PunchthroughAlpha(pBlocks, options, compressionData=None, out=None):
    """cvtt::Kernels::EncodeETC2PunchthroughAlpha, reference ConvectionKernels.h:255 / ConvectionKernels_API.cpp:231-244: pixels whose
    alpha is below options.threshold become the transparent index"""
    return encode("ETC2_PUNCHTHROUGH", pBlocks, options, None, out, _alloc_options(compressionData))


def EncodeETC2Alpha(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeETC2Alpha, reference ConvectionKernels.h:258 / ConvectionKernels_API.cpp:245-255"""
    return encode("ETC2_ALPHA", pBlocks, options, None, out)


def EncodeETC2Alpha11(pBlocks, isSigned, options, out=None):
    """cvtt::Kernels::EncodeETC2Alpha11, reference ConvectionKernels.h:259 / ConvectionKernels_API.cpp:257-268.  pBlocks:
    PixelBlockScalarS16 (int16 [n][16])"""
    return encode("EAC_R11S" if isSigned else "EAC_R11U", pBlocks, options, None, out)


def EncodeBC1(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC1, reference ConvectionKernels.h:243 / ConvectionKernels_API.cpp:86-99"""
    return encode("BC1", pBlocks, options, None, out)


def EncodeBC2(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC2, reference ConvectionKernels.h:244 / ConvectionKernels_API.cpp:101-115"""
    return encode("BC2", pBlocks, options, None, out)


def EncodeBC3(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC3, reference ConvectionKernels.h:245 / ConvectionKernels_API.cpp:117-131"""
    return encode("BC3", pBlocks, options, None, out)


def EncodeBC4U(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC4U, reference ConvectionKernels.h:246 / ConvectionKernels_API.cpp:133-146"""
    return encode("BC4U", pBlocks, options, None, out)


def EncodeBC4S(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC4S (PixelBlockS8 input), reference ConvectionKernels.h:247 / ConvectionKernels_API.cpp:148-164"""
    return encode("BC4S", pBlocks, options, None, out)


def EncodeBC5U(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC5U, reference ConvectionKernels.h:248 / ConvectionKernels_API.cpp:166-180"""
    return encode("BC5U", pBlocks, options, None, out)


def EncodeBC5S(pBlocks, options, out=None):
    """cvtt::Kernels::EncodeBC5S (PixelBlockS8 input), reference ConvectionKernels.h:249 / ConvectionKernels_API.cpp:182-199"""
    return encode("BC5S", pBlocks, options, None, out)


def tiled_block_count(width, height):
    L = _lib()
    L.cvttb200_tiled_block_count.restype = ctypes.c_size_t
    return int(L.cvttb200_tiled_block_count(int(width), int(height)))


def decode(fmt, pBC, out=None):
    """Decodes n 16-byte blocks (numpy array or torch CUDA tensor) of `fmt` ("BC7" -> (n, 16, 4) uint8 PixelBlockU8, "BC6HU" /
    "BC6HS" -> (n, 16, 4) int16 PixelBlockF16 half bits, alpha = 0x3c00)."""
    L = _lib()
    f = FORMATS[fmt] if isinstance(fmt, str) else int(fmt)
    is_bc7 = f == FORMATS["BC7"]
    if _is_torch(pBC):
        import torch
        src = pBC.contiguous()
        n = (src.numel() * src.element_size()) // 16
        if n * 16 != src.numel() * src.element_size():
            raise ValueError("input size is not a whole number of 16-byte blocks")
        if out is None:
            out = torch.empty((n, 16, 4), dtype=torch.uint8 if is_bc7 else torch.int16, device=src.device)
        else:
            _check_out(out, True, n * (64 if is_bc7 else 128), src.device, "uint8" if is_bc7 else "int16")
        stream = torch.cuda.current_stream(src.device).cuda_stream if src.is_cuda else None
        with torch.cuda.device(src.device if src.is_cuda else torch.cuda.current_device()):
            st = L.cvttb200_decode(f, src.data_ptr(), n, out.data_ptr(), stream)
    else:
        src = np.ascontiguousarray(pBC)
        n = (src.size * src.itemsize) // 16
        if n * 16 != src.size * src.itemsize:
            raise ValueError("input size is not a whole number of 16-byte blocks")
        if out is None:
            out = np.empty((n, 16, 4), dtype=np.uint8 if is_bc7 else np.int16)
        else:
            _check_out(out, False, n * (64 if is_bc7 else 128), None, "uint8" if is_bc7 else "int16")
        st = L.cvttb200_decode(f, src.ctypes.data, n, out.ctypes.data, None)
    _check(st)
    return out


def DecodeBC7(pBC, out=None):
    """cvtt::Kernels::DecodeBC7, reference ConvectionKernels.h:275 / ConvectionKernels_API.cpp:288-298"""
    return decode("BC7", pBC, out)


def DecodeBC6HU(pBC, out=None):
    """cvtt::Kernels::DecodeBC6HU, reference ConvectionKernels.h:273 / ConvectionKernels_API.cpp:300-310"""
    return decode("BC6HU", pBC, out)


def DecodeBC6HS(pBC, out=None):
    """cvtt::Kernels::DecodeBC6HS, reference ConvectionKernels.h:274 / ConvectionKernels_API.cpp:312-322"""
    return decode("BC6HS", pBC, out)


def tile_image(image, out=None):
    """image: torch CUDA tensor (H, W, 4) uint8 (RGBA8) or (H, W, 4) int16 / float16 (RGBA16F).  Returns the block array the
    encoders take, (n, 16, 4) of the same dtype, n = tiled_block_count(W, H) (etc2packer order, edges clamped)."""
    import torch
    assert image.is_cuda and image.dim() == 3 and image.shape[2] == 4
    image = image.contiguous()
    h, w = int(image.shape[0]), int(image.shape[1])
    pixel_bytes = 4 * image.element_size()
    n = tiled_block_count(w, h)
    if out is None:
        out = torch.empty((n, 16, 4), dtype=image.dtype, device=image.device)
    else:
        _check_out(out, True, n * 16 * pixel_bytes, image.device, str(image.dtype).replace("torch.", ""))
    L = _lib()
    L.cvttb200_tile_image.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    with torch.cuda.device(image.device):
        _check(L.cvttb200_tile_image(pixel_bytes, image.data_ptr(), w, h, w * pixel_bytes, out.data_ptr(), torch.cuda.current_stream(image.device).cuda_stream))
    return out


def untile_blocks(encoded, width, height):
    """encoded: torch CUDA uint8 tensor (tiled_block_count, blockBytes).  Returns (ceil(H/4) * ceil(W/4), blockBytes) without the
    padding blocks."""
    import torch
    assert encoded.is_cuda and encoded.dim() == 2
    encoded = encoded.contiguous()
    bb = int(encoded.shape[1])
    n = ((height + 3) // 4) * ((width + 3) // 4)
    out = torch.empty((n, bb), dtype=torch.uint8, device=encoded.device)
    L = _lib()
    L.cvttb200_untile_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    with torch.cuda.device(encoded.device):
        _check(L.cvttb200_untile_blocks(encoded.data_ptr(), int(width), int(height), bb, out.data_ptr(), torch.cuda.current_stream(encoded.device).cuda_stream))
    return out


def ktx_header(fmt, width, height):
    """The 68 bytes the reference's sample packer writes in front of the encoded blocks (KTX 1.1 header + imageSize,
    etc2packer/etc2packer.cpp:114-193).  Pure host code."""
    L = _lib()
    buf = ctypes.create_string_buffer(68)
    L.cvttb200_ktx_header.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    _check(L.cvttb200_ktx_header(FORMATS[fmt], int(width), int(height), buf))
    return buf.raw


def pack_ktx(fmt, image, options=None):
    """What the sample packer does with one RGBA8 image (etc2packer/etc2packer.cpp:106-293), on the GPU: block extraction with
    clamped edges, the encode of `fmt`, removal of the padding blocks, KTX header.  image: torch CUDA uint8 (H, W, 4).
    Returns the bytes of the .ktx file.  The two R11 targets take the sample's scalar, the channel sum scaled to 0..2047
    (signed: to 0..1023, as the sample does), etc2packer.cpp:232-237."""
    import torch
    options = options or Options()
    h, w = int(image.shape[0]), int(image.shape[1])
    blocks = tile_image(image)
    if fmt in ("EAC_R11U", "EAC_R11S"):
        total = blocks[..., :3].to(torch.float64).sum(dim=-1) / (255.0 * 3.0)
        scalar = torch.floor(total * (2047.0 if fmt == "EAC_R11U" else 1023.0) + 0.5).to(torch.int16)
        enc = EncodeETC2Alpha11(scalar.contiguous(), fmt == "EAC_R11S", options)
    else:
        enc = encode(fmt, blocks, options)       # AllocETC2Data(options): allocation-time options = the call's, as in the sample
    payload = untile_blocks(enc, w, h)
    return ktx_header(fmt, w, h) + payload.cpu().numpy().tobytes()


def write_ktx(path, fmt, image, options=None):
    with open(path, "wb") as f:
        f.write(pack_ktx(fmt, image, options))

"""CPU ORACLE for the PhantomSDR hot path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings over ``oracle/liboracle.so`` (built from ``oracle/phantom_oracle.c`` by
``oracle/Makefile``) and, when present, ``oracle/_ref/libphantom_ref.so`` (the reference's own
``src/utils/dsp.cpp``, ``src/utils/audioprocessing.cpp`` and ``src/utils.h`` compiled where they
lie under /root/reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package. Nothing under ``phantomsdr_b200/`` does.

Parity pinning status: the reference has no tests or golden vectors for this path and FFTW is not
available, so the FFT call sites are "parity unpinned" against FFTW; everything that compiles
from the reference's sources is pinned bit-for-bit by tests/test_oracle_vs_ref.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_PATH = HERE / "_ref" / "libphantom_ref.so"
REF_FFT_PATH = HERE / "_ref" / "libphantom_ref_fft.so"

USB, LSB, AM, FM = 0, 1, 2, 3  # src/client.h:43
MODE_NAMES = {USB: "USB", LSB: "LSB", AM: "AM", FM: "FM"}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    src_mtime = max((HERE / f).stat().st_mtime for f in ("phantom_oracle.c", "fft_generic.inc", "Makefile"))
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src_mtime:
        subprocess.check_call(["make", "-s", "-C", str(HERE), str(LIB_PATH)])
    if Path("/root/reference/src/utils").is_dir():
        shim_mtime = max((HERE / f).stat().st_mtime for f in ("ref_shim.cpp", "ref_shim_fft.cpp", "Makefile"))
        if force or not REF_PATH.exists() or not REF_FFT_PATH.exists() or \
                min(REF_PATH.stat().st_mtime, REF_FFT_PATH.stat().st_mtime) < shim_mtime:
            subprocess.check_call(["make", "-s", "-C", str(HERE), "ref"])


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, sz, i, f, d = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_double
    sig = {
        "orc_audio_fft_size": (i, [i, i, i]),
        "orc_downsample_levels": (i, [i, i]),
        "orc_skip_num": (i, [i, i]),
        "orc_hann_window": (None, [_f32p, i]),
        "orc_convert_u8": (None, [vp, _f32p, sz]),
        "orc_convert_s8": (None, [vp, _f32p, sz]),
        "orc_convert_u16": (None, [vp, _f32p, sz]),
        "orc_convert_s16": (None, [vp, _f32p, sz]),
        "orc_quantize_one": (C.c_int8, [f, i]),
        "orc_fft_create": (vp, [sz, i, i]),
        "orc_fft_set_output_additional_size": (None, [vp, sz]),
        "orc_fft_plan_c2c": (i, [vp]),
        "orc_fft_plan_r2c": (i, [vp]),
        "orc_fft_destroy": (None, [vp]),
        "orc_fft_window": (vp, [vp]),
        "orc_fft_input": (vp, [vp]),
        "orc_fft_output": (vp, [vp]),
        "orc_fft_power": (vp, [vp]),
        "orc_fft_quantized": (vp, [vp]),
        "orc_fft_size_log2": (i, [vp]),
        "orc_fft_load_real_input": (i, [vp, _f32p, _f32p]),
        "orc_fft_load_complex_input": (i, [vp, _f32p, _f32p]),
        "orc_fft_transform": (i, [vp]),
        "orc_fft_quantize": (i, [vp]),
        "orc_fft_execute": (i, [vp]),
        "orc_fft_wrap_copy": (None, [vp, sz]),
        "orc_fft_shadow_f64": (None, [vp, _f64p]),
        "orc_dft_f32": (None, [_f32p, _f32p, C.c_long, i]),
        "orc_dft_f64": (None, [_f64p, _f64p, C.c_long, i]),
        "orc_dc_create": (vp, [i]),
        "orc_dc_destroy": (None, [vp]),
        "orc_dc_remove": (None, [vp, _f32p, i]),
        "orc_agc_create": (vp, [f, f, f, f, f]),
        "orc_agc_destroy": (None, [vp]),
        "orc_agc_attack": (f, [vp]),
        "orc_agc_release": (f, [vp]),
        "orc_agc_lookahead": (sz, [vp]),
        "orc_agc_gain": (f, [vp]),
        "orc_agc_process": (None, [vp, _f32p, sz]),
        "orc_agc_reset": (None, [vp]),
        "orc_polar_discriminator_fm": (None, [_f32p, f, f, _f32p, sz]),
        "orc_am_demod": (None, [_f32p, _f32p, sz]),
        "orc_float_to_int16": (None, [_f32p, _i32p, f, sz]),
        "orc_client_create": (vp, [i, i, i, i]),
        "orc_client_destroy": (None, [vp]),
        "orc_client_set_audio_range": (None, [vp, i, d, i]),
        "orc_client_on_window_message": (i, [vp, i, d, i]),
        "orc_client_set_audio_demodulation": (None, [vp, i]),
        "orc_client_on_demodulation_message": (None, [vp, i]),
        "orc_client_send_audio": (i, [vp, vp, sz, vp, vp, vp]),
        "orc_signal_slice_offset": (sz, [i, sz, i]),
        "orc_waterfall_level_offset": (sz, [i, sz]),
        "orc_clients_send_audio": (None, [vp, i, vp, sz, i, sz, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


_ref = None


def ref():
    """The reference's own compiled helpers (oracle/_ref), or None if not built."""
    global _ref
    if _ref is not None:
        return _ref
    if not REF_PATH.exists():
        return None
    R = C.CDLL(str(REF_PATH))
    vp, sz, i, f = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    sig = {
        "ref_build_hann_window": (None, [_f32p, i]),
        "ref_polar_discriminator_fm": (None, [_f32p, f, f, _f32p, sz]),
        "ref_dsp_negate_float": (None, [_f32p, sz]),
        "ref_dsp_negate_complex": (None, [_f32p, sz]),
        "ref_dsp_add_float": (None, [_f32p, _f32p, sz]),
        "ref_dsp_add_complex": (None, [_f32p, _f32p, sz]),
        "ref_dsp_am_demod": (None, [_f32p, _f32p, sz]),
        "ref_dsp_float_to_int16": (None, [_f32p, _i32p, f, sz]),
        "ref_agc_create": (vp, [f, f, f, f, f]),
        "ref_agc_destroy": (None, [vp]),
        "ref_agc_process": (None, [vp, _f32p, sz]),
        "ref_agc_reset": (None, [vp]),
        "ref_dc_create": (vp, [i]),
        "ref_dc_destroy": (None, [vp]),
        "ref_dc_remove": (None, [vp, _f32p, i]),
        "ref_slice_power": (f, [_f32p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(R, name)
        fn.restype = res
        fn.argtypes = args
    _ref = R
    return R


_ref_fft = None


def ref_fft():
    """The reference's own src/fft_impl.cpp (class FFTW) and src/signal.cpp (AudioClient::send_audio) compiled against
    stand-in third-party headers (oracle/ref_shim_fft.cpp), or None if not built. The DFT behind fftwf_execute is the
    oracle's own (FFTW3f is absent), so everything around the transforms compares bit for bit."""
    global _ref_fft
    if _ref_fft is not None:
        return _ref_fft
    if not REF_FFT_PATH.exists():
        return None
    R = C.CDLL(str(REF_FFT_PATH))
    vp, sz, i, f, d = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_double
    sig = {
        "ref_set_dft": (None, [vp]),
        "ref_fftw_create": (vp, [sz, i, i, sz, i]),
        "ref_fftw_destroy": (None, [vp]),
        "ref_fftw_load_real": (None, [vp, _f32p, _f32p]),
        "ref_fftw_load_complex": (None, [vp, _f32p, _f32p]),
        "ref_fftw_execute": (None, [vp]),
        "ref_fftw_input": (vp, [vp]),
        "ref_fftw_output": (vp, [vp]),
        "ref_fftw_quantized": (vp, [vp]),
        "ref_audio_create": (vp, [i, i, i, i]),
        "ref_audio_destroy": (None, [vp]),
        "ref_audio_set_range": (None, [vp, i, d, i]),
        "ref_audio_set_demodulation": (None, [vp, i]),
        "ref_audio_on_demodulation_message": (None, [vp, C.c_char_p]),
        "ref_audio_on_window_message": (i, [vp, i, d, i]),
        "ref_audio_send": (i, [vp, vp, sz, _i32p, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(R, name)
        fn.restype = res
        fn.argtypes = args
    R.ref_set_dft(C.cast(lib().orc_dft_f32, vp))  # both sides transform with the same arithmetic
    _ref_fft = R
    return R


def _view(ptr: int, shape, dtype) -> np.ndarray:
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


# ------------------------------------------------------------------ derived sizes
def audio_fft_size(audio_max_sps: int, fft_size: int, sps: int) -> int:
    return lib().orc_audio_fft_size(audio_max_sps, fft_size, sps)


def downsample_levels(fft_result_size: int, min_waterfall_fft: int = 1024) -> int:
    return lib().orc_downsample_levels(fft_result_size, min_waterfall_fft)


def skip_num(sps: int, fft_size: int) -> int:
    return lib().orc_skip_num(sps, fft_size)


def hann_window(n: int) -> np.ndarray:
    w = np.empty(n, np.float32)
    lib().orc_hann_window(w, n)
    return w


def convert(raw: np.ndarray) -> np.ndarray:
    """samplereader.cpp conversion for u8/s8/u16/s16 numpy arrays."""
    raw = np.ascontiguousarray(raw)
    out = np.empty(raw.size, np.float32)
    fn = {"uint8": "orc_convert_u8", "int8": "orc_convert_s8", "uint16": "orc_convert_u16",
          "int16": "orc_convert_s16"}[raw.dtype.name]
    getattr(lib(), fn)(raw.ctypes.data, out, raw.size)
    return out


def slice_offset(l: int, fft_size: int, is_real: bool) -> int:
    return lib().orc_signal_slice_offset(l, fft_size, int(is_real))


def level_offset(level: int, fft_result_size: int) -> int:
    return lib().orc_waterfall_level_offset(level, fft_result_size)


def pyramid_size(fft_result_size: int, levels: int) -> int:
    return sum(fft_result_size >> i for i in range(levels))


# ------------------------------------------------------------------ FFT backend
class OracleFFT:
    """Restatement of class FFTW (src/fft.h:93-105, src/fft_impl.cpp)."""

    def __init__(self, size: int, downsample_levels: int, brightness_offset: int = 0):
        self.L = lib()
        self.size = size
        self.levels = downsample_levels
        self.h = self.L.orc_fft_create(size, downsample_levels, brightness_offset)
        self.additional = 0
        self.is_real = None

    def set_output_additional_size(self, n: int):
        self.additional = n
        self.L.orc_fft_set_output_additional_size(self.h, n)

    def plan_c2c(self):
        self.is_real = False
        self.L.orc_fft_plan_c2c(self.h)

    def plan_r2c(self):
        self.is_real = True
        self.L.orc_fft_plan_r2c(self.h)

    @property
    def result_size(self):
        return self.size // 2 if self.is_real else self.size

    @property
    def window(self):
        return _view(self.L.orc_fft_window(self.h), (self.size,), np.float32)

    @property
    def inbuf(self):
        n = self.size if self.is_real else 2 * self.size
        return _view(self.L.orc_fft_input(self.h), (n,), np.float32)

    @property
    def outbuf(self):
        """float32 view of outbuf: complex interleaved, R (+additional) or N/2+1 bins."""
        n = self.size + 2 if self.is_real else 2 * (self.size + self.additional)
        return _view(self.L.orc_fft_output(self.h), (n,), np.float32)

    @property
    def spectrum(self):
        return self.outbuf.view(np.complex64)

    @property
    def powerbuf(self):
        return _view(self.L.orc_fft_power(self.h), (pyramid_size(self.result_size, self.levels),), np.float32)

    @property
    def quantized(self):
        return _view(self.L.orc_fft_quantized(self.h), (pyramid_size(self.result_size, self.levels),), np.int8)

    def load_real_input(self, a1, a2):
        self.L.orc_fft_load_real_input(self.h, np.ascontiguousarray(a1, np.float32), np.ascontiguousarray(a2, np.float32))

    def load_complex_input(self, a1, a2):
        a1 = np.ascontiguousarray(a1).view(np.float32)
        a2 = np.ascontiguousarray(a2).view(np.float32)
        self.L.orc_fft_load_complex_input(self.h, a1, a2)

    def transform(self):
        self.L.orc_fft_transform(self.h)

    def quantize(self):
        self.L.orc_fft_quantize(self.h)

    def execute(self):
        self.L.orc_fft_execute(self.h)

    def wrap_copy(self, n):
        self.L.orc_fft_wrap_copy(self.h, n)

    def shadow_f64(self) -> np.ndarray:
        n = self.size + 2 if self.is_real else 2 * self.size
        out = np.empty(n, np.float64)
        self.L.orc_fft_shadow_f64(self.h, out)
        return out.view(np.complex128)

    def requantize_from(self, spectrum_normalised: np.ndarray) -> np.ndarray:
        """Run ONLY the integer/bit-trick quantiser + pyramid (fft_impl.cpp:146-173) on a given
        already-normalised spectrum (e.g. the CUDA path's output), by undoing the exact
        power-of-two normalisation first. Returns the int8 pyramid."""
        raw = np.ascontiguousarray(spectrum_normalised).view(np.float32).copy()
        nfl = self.size + 2 if self.is_real else 2 * self.size
        out = self.outbuf
        out[:nfl] = raw[:nfl] * np.float32(self.size)  # exact: size is a power of two
        self.quantize()
        return self.quantized.copy()

    def close(self):
        if self.h:
            self.L.orc_fft_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dft(x: np.ndarray, sign: int) -> np.ndarray:
    x = np.ascontiguousarray(x)
    if x.dtype == np.complex64:
        out = np.empty_like(x)
        lib().orc_dft_f32(x.view(np.float32), out.view(np.float32), x.size, sign)
        return out
    x = x.astype(np.complex128)
    out = np.empty_like(x)
    lib().orc_dft_f64(x.view(np.float64), out.view(np.float64), x.size, sign)
    return out


# ------------------------------------------------------------------ per-client chain
class OracleDC:
    def __init__(self, delay: int):
        self.L = lib()
        self.h = self.L.orc_dc_create(delay)

    def remove(self, arr: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(arr, np.float32).copy()
        self.L.orc_dc_remove(self.h, a, a.size)
        return a

    def __del__(self):
        try:
            self.L.orc_dc_destroy(self.h)
        except Exception:
            pass


class OracleAGC:
    def __init__(self, level=0.2, attack_ms=50.0, release_ms=300.0, lookahead_ms=200.0, sr=12000.0):
        self.L = lib()
        self.h = self.L.orc_agc_create(level, attack_ms, release_ms, lookahead_ms, sr)

    @property
    def attack(self):
        return self.L.orc_agc_attack(self.h)

    @property
    def release(self):
        return self.L.orc_agc_release(self.h)

    @property
    def lookahead(self):
        return self.L.orc_agc_lookahead(self.h)

    def process(self, arr: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(arr, np.float32).copy()
        self.L.orc_agc_process(self.h, a, a.size)
        return a

    def reset(self):
        self.L.orc_agc_reset(self.h)

    def __del__(self):
        try:
            self.L.orc_agc_destroy(self.h)
        except Exception:
            pass


class OracleClient:
    """Restatement of AudioClient (src/signal.cpp)."""

    def __init__(self, is_real: bool, audio_fft_size: int, audio_max_sps: int, fft_result_size: int):
        self.L = lib()
        self.n = audio_fft_size
        self.is_real = bool(is_real)
        self.h = self.L.orc_client_create(int(is_real), audio_fft_size, audio_max_sps, fft_result_size)
        self.l = self.r = 0
        self.mid = 0.0
        self.mode = USB

    def set_audio_range(self, l: int, m: float, r: int):
        self.l, self.mid, self.r = l, m, r
        self.L.orc_client_set_audio_range(self.h, l, m, r)

    def on_window_message(self, l, m, r) -> bool:
        ok = bool(self.L.orc_client_on_window_message(self.h, l, m, r))
        if ok:
            self.l, self.mid, self.r = l, m, r
        return ok

    def set_audio_demodulation(self, mode: int):
        self.mode = mode
        self.L.orc_client_set_audio_demodulation(self.h, mode)

    def on_demodulation_message(self, mode: int):
        self.mode = mode
        self.L.orc_client_on_demodulation_message(self.h, mode)

    def send_audio(self, spectrum: np.ndarray, fft_size: int, frame_num: int):
        """spectrum: the whole outbuf (complex64, with wrap tail for IQ). Forms the slice pointer as
        websocket.cpp:182 does. Returns (valid, pcm int32[n/2], pwr float, audio_pre_dc float32[n/2])."""
        spec = np.ascontiguousarray(spectrum).view(np.float32)
        off = slice_offset(self.l, fft_size, self.is_real)
        pcm = np.zeros(self.n // 2, np.int32)
        pwr = np.zeros(1, np.float32)
        pre = np.zeros(self.n // 2, np.float32)
        ok = self.L.orc_client_send_audio(self.h, spec.ctypes.data + 8 * off, frame_num, pcm.ctypes.data,
                                          pwr.ctypes.data, pre.ctypes.data)
        return bool(ok), pcm, float(pwr[0]), pre

    def __del__(self):
        try:
            self.L.orc_client_destroy(self.h)
        except Exception:
            pass


def clients_send_audio(clients, spectrum: np.ndarray, fft_size: int, is_real: bool, frame_num: int):
    """OpenMP-parallel batch (timed CPU baseline). Returns (pcm[nc, n/2], pwr[nc], valid[nc])."""
    nc = len(clients)
    half = clients[0].n // 2
    arr = (C.c_void_p * nc)(*[c.h for c in clients])
    spec = np.ascontiguousarray(spectrum).view(np.float32)
    pcm = np.zeros((nc, half), np.int32)
    pwr = np.zeros(nc, np.float32)
    valid = np.zeros(nc, np.uint8)
    lib().orc_clients_send_audio(arr, nc, spec.ctypes.data, fft_size, int(is_real), frame_num, pcm.ctypes.data,
                                 pwr.ctypes.data, valid.ctypes.data)
    return pcm, pwr, valid

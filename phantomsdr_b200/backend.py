"""Host-side mirror of the reference's FFT-backend interface and client slots, over the C ABI.

``B200FFT`` has the method names, argument meaning and call order of ``class FFT``
(reference src/fft.h:33-63): ``malloc/free``, ``set_output_additional_size``,
``plan_c2c/plan_r2c``, ``load_real_input/load_complex_input``, ``execute``,
``get_output_buffer/get_quantized_buffer``. The audio clients that the reference runs one
``AudioClient::send_audio`` task at a time (src/signal.cpp:102-298, src/websocket.cpp:156-185) are
a batched slot table here (``clients_*``). Every method is a thin call into
``libphantomsdr_b200.so``; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import B200Error, check  # noqa: F401

FORWARD, BACKWARD = 0, 1
FMT_F32, FMT_U8, FMT_S8, FMT_U16, FMT_S16 = 0, 1, 2, 3, 4
OPT_RELOAD_BOTH, OPT_HOST_MIRROR, OPT_INPUT_FORMAT, OPT_STAGE_MASK, OPT_FUSED_PYRAMID, OPT_TMA, OPT_TAIL_PIPELINE, OPT_PEER_STORES = 1, 2, 3, 4, 5, 6, 7, 8
OPT_PACKED_MATH, OPT_FWD_LANES, OPT_FWD_SUB_FRAMES, OPT_PASS1_ORDER = 9, 10, 11, 12
OPT_PYRAMID_LAG = 13
OPT_STREAM_GRID, OPT_STREAM_LAG1, OPT_STREAM_LAG2, OPT_STREAM_RING = 14, 15, 16, 17
OPT_PCM16 = 18
OPT_DEMOD_CHUNK = 19
OPT_CLIENT_STAGE_MASK = 20
OPT_FWD_SMS = 21
OPT_PASS1_SPLIT = 22
OPT_DEMOD_GENERIC = 23
OPT_TAIL_SMEM_KB = 24
OPT_R2C_SPLIT_KERNEL = 25
_PUBLIC_OPTIONS = {OPT_RELOAD_BOTH, OPT_HOST_MIRROR, OPT_INPUT_FORMAT, OPT_PEER_STORES, OPT_PCM16}  # include/phantomsdr_b200.h
_FMT_OF_DTYPE = {"float32": FMT_F32, "uint8": FMT_U8, "int8": FMT_S8, "uint16": FMT_U16, "int16": FMT_S16}


def device_count() -> int:
    return _ffi.lib().b200_device_count()


def _host_view(ptr: int, count: int, dtype) -> np.ndarray:
    nbytes = count * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class _DevArray:
    """Zero-copy handle on engine-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


class B200FFT:
    """FFT backend on one B200. Mirrors ``class FFT`` / ``class cuFFT`` (src/fft.h:33-63,127-145)."""

    def __init__(self, size: int, nthreads: int = 1, downsample_levels: int = 1, brightness_offset: int = 0,
                 device: int = 0):
        self.L = _ffi.lib()
        self.size = int(size)
        self.levels = int(downsample_levels)
        self.device = device
        self.additional = 0
        self.is_real: Optional[bool] = None
        self._bufs = {}
        h = C.c_void_p()
        check(self.L.b200_engine_create(C.byref(h), self.size, nthreads, downsample_levels, brightness_offset, device))
        self.h = h
        self.n_audio = 0
        self.max_clients = 0

    # ---- class FFT surface -------------------------------------------------------------------
    def set_output_additional_size(self, n: int) -> None:
        check(self.L.b200_set_output_additional_size(self.h, n))
        self.additional = int(n)

    def malloc(self, nfloats: int) -> np.ndarray:
        p = self.L.b200_malloc(self.h, nfloats)
        if not p:
            raise B200Error(-12, self.L.b200_last_error().decode())
        a = _host_view(p, nfloats, np.float32)
        self._bufs[a.ctypes.data] = p
        return a

    def free(self, buf: np.ndarray) -> None:
        p = self._bufs.pop(buf.ctypes.data, None)
        if p:
            self.L.b200_free(self.h, p)

    def plan_c2c(self, direction: int = FORWARD, options: int = 0) -> int:
        check(self.L.b200_plan_c2c(self.h, direction, options))
        self.is_real = False
        return 0

    def plan_r2c(self, options: int = 0) -> int:
        check(self.L.b200_plan_r2c(self.h, options))
        self.is_real = True
        return 0

    @property
    def result_size(self) -> int:
        return self.size // 2 if self.is_real else self.size

    def get_output_buffer(self) -> np.ndarray:
        """float32 view of outbuf (complex interleaved), stable for the engine's lifetime."""
        n = self.size + 2 if self.is_real else 2 * (self.size + self.additional)
        return _host_view(self.L.b200_get_output_buffer(self.h), n, np.float32)

    def get_quantized_buffer(self) -> np.ndarray:
        return _host_view(self.L.b200_get_quantized_buffer(self.h), self.L.b200_pyramid_bytes(self.h), np.int8)

    def load_real_input(self, a1: np.ndarray, a2: np.ndarray) -> int:
        check(self.L.b200_load_real_input(self.h, _ptr(a1), _ptr(a2)))
        return 0

    def load_complex_input(self, a1: np.ndarray, a2: np.ndarray) -> int:
        check(self.L.b200_load_complex_input(self.h, _ptr(a1), _ptr(a2)))
        return 0

    def load_raw_input(self, a1: np.ndarray, a2: np.ndarray) -> int:
        """SampleConverter<T> fused on the GPU (src/samplereader.cpp:29-66); dtype picks the format."""
        check(self.L.b200_set_option(self.h, OPT_INPUT_FORMAT, _FMT_OF_DTYPE[a1.dtype.name]))
        check(self.L.b200_load_raw_input(self.h, _ptr(a1), _ptr(a2)))
        return 0

    def execute(self) -> int:
        check(self.L.b200_execute(self.h))
        return 0

    def set_option(self, option: int, value: int) -> None:
        """Engine options go through b200_set_option, tuning knobs (phantomsdr_b200_debug.h) through b200_debug_option."""
        if option in _PUBLIC_OPTIONS:
            check(self.L.b200_set_option(self.h, option, value))
        else:
            check(self.L.b200_debug_option(self.h, option, value))

    def set_waterfall_cadence(self, skip_num: int) -> None:
        """Pyramids only for frames with frame_num % skip_num == 0 (src/fft.cpp:33,102-104)."""
        check(self.L.b200_set_waterfall_cadence(self.h, skip_num))

    def set_frame_number(self, frame_num: int) -> None:
        check(self.L.b200_set_frame_number(self.h, frame_num))

    # ---- device-resident streaming form ------------------------------------------------------
    def set_hop_ring(self, nhops: int) -> None:
        check(self.L.b200_set_hop_ring(self.h, nhops))
        self.nhops = nhops

    def set_batch_frames(self, frames: int) -> None:
        check(self.L.b200_set_batch_frames(self.h, frames))

    def set_pipeline(self, banks: int) -> None:
        check(self.L.b200_set_pipeline(self.h, banks))

    def select_bank(self, bank: int) -> None:
        check(self.L.b200_select_bank(self.h, bank))

    def bank_acquire(self) -> None:
        check(self.L.b200_bank_acquire(self.h))

    def join_streams(self) -> None:
        check(self.L.b200_join_streams(self.h))

    def client_stream_wait_event(self, cuda_event: int) -> None:
        check(self.L.b200_client_stream_wait_event(self.h, cuda_event))

    @property
    def hop_floats(self) -> int:
        return self.L.b200_hop_floats(self.h)

    @property
    def spectrum_bins(self) -> int:
        return self.L.b200_spectrum_bins(self.h)

    @property
    def spectrum_stride(self) -> int:
        return self.L.b200_spectrum_stride(self.h)

    @property
    def pyramid_bytes(self) -> int:
        return self.L.b200_pyramid_bytes(self.h)

    @property
    def pyramid_stride(self) -> int:
        return self.L.b200_pyramid_stride(self.h)

    def device_hop_ring(self, nhops: int) -> _DevArray:
        return _DevArray(self.L.b200_device_hop_ring(self.h), (nhops, self.hop_floats), "<f4", self)

    def device_spectrum(self, frames: int = 1) -> _DevArray:
        """float32 [frames][stride][2] view of the device spectrum."""
        return _DevArray(self.L.b200_device_spectrum(self.h), (frames, self.spectrum_stride, 2), "<f4", self)

    def device_quantized(self, frames: int = 1) -> _DevArray:
        return _DevArray(self.L.b200_device_quantized(self.h), (frames, self.pyramid_stride), "|i1", self)

    def execute_device(self, hop_index: int, nframes: int = 1) -> None:
        check(self.L.b200_execute_device_batch(self.h, hop_index, nframes))

    def sync(self) -> None:
        check(self.L.b200_sync(self.h))

    @property
    def stream(self) -> int:
        return self.L.b200_stream(self.h) or 0

    def set_stream(self, cuda_stream: int) -> None:
        """Enqueue on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = own."""
        check(self.L.b200_set_stream(self.h, cuda_stream or None))

    def bind_spectrum(self, dev_ptr: int) -> None:
        check(self.L.b200_bind_spectrum(self.h, dev_ptr))

    def set_peer_spectra(self, ptrs: Sequence[int]) -> None:
        arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs)
        check(self.L.b200_set_peer_spectra(self.h, len(ptrs), arr))

    def set_peer_ranges(self, peer: int, lo0: int, hi0: int, lo1: int = 0, hi1: int = 0) -> None:
        check(self.L.b200_set_peer_ranges(self.h, peer, lo0, hi0, lo1, hi1))

    def push_peers(self, nframes: int) -> None:
        check(self.L.b200_push_peers(self.h, nframes))

    def pull_spectrum(self, remote_spectrum: int, nframes: int, lo0: int, hi0: int, lo1: int = 0, hi1: int = 0) -> None:
        check(self.L.b200_pull_spectrum(self.h, remote_spectrum, nframes, lo0, hi0, lo1, hi1))

    @property
    def spectrum_base(self) -> int:
        return self.L.b200_device_spectrum_base(self.h)

    @property
    def spectrum_offset(self) -> int:
        return self.L.b200_device_spectrum_offset(self.h)

    @property
    def flag_buffer(self) -> int:
        return self.L.b200_flag_buffer(self.h)

    # stream selectors of the flag calls: 0 forward, 1 client, 2 copy stream
    def enqueue_signal(self, client_stream, flag_ptrs: Sequence[int], value: int) -> None:
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        check(self.L.b200_enqueue_signal(self.h, int(client_stream), arr, len(flag_ptrs), value))

    def enqueue_wait(self, client_stream, flag_ptrs: Sequence[int], min_value: int, timeout_ms: int = 2000) -> None:
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        check(self.L.b200_enqueue_wait(self.h, int(client_stream), arr, len(flag_ptrs), min_value, timeout_ms))

    @property
    def flag_error(self) -> int:
        return self.L.b200_flag_error(self.h)

    def ipc_export(self, dev_ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        check(self.L.b200_ipc_export(self.h, dev_ptr, buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        check(self.L.b200_ipc_open(self.h, buf, C.byref(out)))
        return out.value

    def ipc_close(self, dev_ptr: int) -> None:
        check(self.L.b200_ipc_close(self.h, dev_ptr))

    @property
    def launch_count(self) -> int:
        return self.L.b200_launch_count(self.h)

    # ---- signal slot (AudioClient) -----------------------------------------------------------
    def clients_create(self, max_clients: int, audio_fft_size: int, audio_max_sps: int) -> None:
        check(self.L.b200_clients_create(self.h, max_clients, audio_fft_size, audio_max_sps))
        self.max_clients = max_clients
        self.n_audio = audio_fft_size

    def client_open(self, slot: int, l: int, audio_mid: float, r: int, demodulation: int) -> None:
        check(self.L.b200_client_open(self.h, slot, l, audio_mid, r, demodulation))

    def client_set_window(self, slot: int, l: int, audio_mid: float, r: int) -> bool:
        """AudioClient::on_window_message: returns False where the reference ignores the message."""
        rc = self.L.b200_client_set_window(self.h, slot, l, audio_mid, r)
        if rc == -22:
            return False
        check(rc)
        return True

    def client_set_demodulation(self, slot: int, demodulation: int) -> None:
        check(self.L.b200_client_set_demodulation(self.h, slot, demodulation))

    def client_close(self, slot: int) -> None:
        check(self.L.b200_client_close(self.h, slot))

    def clients_execute(self, frame_num: int):
        """One signal_loop pass. Returns (pcm int32 [max_clients, n/2], pwr float32 [max_clients], valid uint8)."""
        h = self.n_audio // 2
        pcm = np.zeros((self.max_clients, h), np.int32)
        pwr = np.zeros(self.max_clients, np.float32)
        valid = np.zeros(self.max_clients, np.uint8)
        check(self.L.b200_clients_execute(self.h, frame_num, _ptr(pcm), _ptr(pwr), _ptr(valid)))
        return pcm, pwr, valid

    def clients_execute_device(self, frame_num: int, nframes: int = 1) -> None:
        check(self.L.b200_clients_execute_device(self.h, frame_num, nframes))

    def clients_fetch(self, frame: int = 0, out=None):
        h = self.n_audio // 2
        if out is None:
            out = (np.zeros((self.max_clients, h), np.int32), np.zeros(self.max_clients, np.float32),
                   np.zeros(self.max_clients, np.uint8))
        check(self.L.b200_clients_fetch(self.h, frame, _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    def clients_fetch_async(self, slot: int, nframes: int, pcm=None, pwr=None, valid=None) -> None:
        """Enqueue the D2H copies of the last client batch into page-locked host arrays; see clients_fetch_wait."""
        check(self.L.b200_clients_fetch_async(self.h, slot, nframes, _ptr(pcm), _ptr(pwr), _ptr(valid)))

    def clients_fetch_wait(self, slot: int) -> None:
        check(self.L.b200_clients_fetch_wait(self.h, slot))

    def clients_read_pre_dc(self) -> np.ndarray:
        out = np.zeros((self.max_clients, self.n_audio // 2), np.float32)
        check(self.L.b200_clients_read_pre_dc(self.h, _ptr(out)))
        return out

    # ---- pipelined host-block streaming -----------------------------------------------------
    def stream_prime(self, older_half: np.ndarray) -> None:
        check(self.L.b200_stream_prime(self.h, _ptr(older_half)))

    def submit_block(self, new_halves, frame_num0: int, pcm=None, pwr=None, valid=None, pyramid=None) -> None:
        """new_halves: sequence of host arrays (one per frame). Outputs: pinned host arrays or None."""
        n = len(new_halves)
        arr = (C.c_void_p * n)(*[_ptr(a) for a in new_halves])
        check(self.L.b200_submit_block(self.h, arr, n, frame_num0, _ptr(pcm), _ptr(pwr), _ptr(valid), _ptr(pyramid)))

    def wait_block(self) -> None:
        check(self.L.b200_wait_block(self.h))

    def pinned(self, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """Page-locked host array (through FFT::malloc) viewed as dtype."""
        a = self.malloc((nbytes + 3) // 4)
        return a.view(np.uint8)[:nbytes].view(dtype)

    # ---- waterfall slot ----------------------------------------------------------------------
    def waterfall_gather(self, levels: Sequence[int], ls: Sequence[int], rs: Sequence[int]):
        """N x send_waterfall (src/waterfall.cpp:44-51): returns one int8 array per client."""
        n = len(levels)
        lv = np.asarray(levels, np.int32)
        l = np.asarray(ls, np.int32)
        r = np.asarray(rs, np.int32)
        lens = (r - l).astype(np.int64)
        offs = np.zeros(n, np.uint64)
        if n:
            offs[1:] = np.cumsum(lens)[:-1]
        out = np.zeros(int(lens.sum()) if n else 0, np.int8)
        check(self.L.b200_waterfall_gather(self.h, n, _ptr(lv), _ptr(l), _ptr(r), _ptr(offs), _ptr(out)))
        return [out[int(offs[i]): int(offs[i] + lens[i])] for i in range(n)]

    def close(self) -> None:
        if getattr(self, "h", None):
            for p in list(self._bufs.values()):
                self.L.b200_free(self.h, p)
            self._bufs.clear()
            self.L.b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def quant_table(power_offset: int):
    """(lo, hi, base) of the table-driven waterfall quantiser for one power offset (host-only, b200_quant_table)."""
    lo = np.zeros(2048, np.uint32)
    hi = np.zeros(2048, np.uint32)
    base = np.zeros(2048, np.uint8)
    check(_ffi.lib().b200_quant_table(int(power_offset), lo.ctypes.data, hi.ctypes.data, base.ctypes.data))
    return lo, hi, base

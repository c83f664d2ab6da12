"""ctypes binding of include/phantomsdr_b200.h. Loads phantomsdr_b200/lib/libphantomsdr_b200.so
(built in-tree by phantomsdr_b200/build.py). Fails loudly when the library is missing - there is
no CPU or PyTorch fallback for the compute path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libphantomsdr_b200.so"
HEADER = PKG.parent / "include" / "phantomsdr_b200.h"
DEBUG_HEADER = PKG.parent / "include" / "phantomsdr_b200_debug.h"  # tuning knobs, profiling hooks

_vp, _sz, _i, _d, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_uint64
_pp = C.POINTER(C.c_void_p)

SIGNATURES = {
    "b200_abi_version": (_i, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (_i, []),
    "b200_engine_create": (_i, [_pp, _sz, _i, _i, _i, _i]),
    "b200_engine_destroy": (None, [_vp]),
    "b200_set_output_additional_size": (_i, [_vp, _sz]),
    "b200_malloc": (_vp, [_vp, _sz]),
    "b200_free": (None, [_vp, _vp]),
    "b200_plan_c2c": (_i, [_vp, _i, _i]),
    "b200_plan_r2c": (_i, [_vp, _i]),
    "b200_get_output_buffer": (_vp, [_vp]),
    "b200_get_quantized_buffer": (_vp, [_vp]),
    "b200_load_real_input": (_i, [_vp, _vp, _vp]),
    "b200_load_complex_input": (_i, [_vp, _vp, _vp]),
    "b200_execute": (_i, [_vp]),
    "b200_set_option": (_i, [_vp, _i, _i]),
    "b200_debug_option": (_i, [_vp, _i, _i]),
    "b200_load_raw_input": (_i, [_vp, _vp, _vp]),
    "b200_device_spectrum": (_vp, [_vp]),
    "b200_device_quantized": (_vp, [_vp]),
    "b200_device_hop_ring": (_vp, [_vp]),
    "b200_hop_floats": (_sz, [_vp]),
    "b200_spectrum_bins": (_sz, [_vp]),
    "b200_pyramid_bytes": (_sz, [_vp]),
    "b200_set_hop_ring": (_i, [_vp, _sz]),
    "b200_execute_device": (_i, [_vp, _sz]),
    "b200_set_batch_frames": (_i, [_vp, _i]),
    "b200_execute_device_batch": (_i, [_vp, _sz, _i]),
    "b200_spectrum_stride": (_sz, [_vp]),
    "b200_pyramid_stride": (_sz, [_vp]),
    "b200_set_pipeline": (_i, [_vp, _i]),
    "b200_select_bank": (_i, [_vp, _i]),
    "b200_bank_acquire": (_i, [_vp]),
    "b200_join_streams": (_i, [_vp]),
    "b200_client_stream_wait_event": (_i, [_vp, _vp]),
    "b200_sync": (_i, [_vp]),
    "b200_stream": (_vp, [_vp]),
    "b200_set_stream": (_i, [_vp, _vp]),
    "b200_bind_spectrum": (_i, [_vp, _vp]),
    "b200_set_peer_spectra": (_i, [_vp, _i, _pp]),
    "b200_set_peer_ranges": (_i, [_vp, _i, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "b200_device_spectrum_base": (_vp, [_vp]),
    "b200_device_spectrum_offset": (_sz, [_vp]),
    "b200_push_peers": (_i, [_vp, _i]),
    "b200_pull_spectrum": (_i, [_vp, _vp, _i, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "b200_flag_buffer": (_vp, [_vp]),
    "b200_enqueue_signal": (_i, [_vp, _i, _pp, _i, _u64]),
    "b200_enqueue_wait": (_i, [_vp, _i, _pp, _i, _u64, _i]),
    "b200_flag_error": (_i, [_vp]),
    "b200_ipc_export": (_i, [_vp, _vp, _vp]),
    "b200_ipc_open": (_i, [_vp, _vp, _pp]),
    "b200_ipc_close": (_i, [_vp, _vp]),
    "b200_clients_create": (_i, [_vp, _i, _i, _i]),
    "b200_client_open": (_i, [_vp, _i, _i, _d, _i, _i]),
    "b200_client_set_window": (_i, [_vp, _i, _i, _d, _i]),
    "b200_client_set_demodulation": (_i, [_vp, _i, _i]),
    "b200_client_close": (_i, [_vp, _i]),
    "b200_clients_execute": (_i, [_vp, _u64, _vp, _vp, _vp]),
    "b200_clients_execute_device": (_i, [_vp, _u64, _i]),
    "b200_device_pcm": (_vp, [_vp]),
    "b200_device_pwr": (_vp, [_vp]),
    "b200_device_valid": (_vp, [_vp]),
    "b200_clients_fetch": (_i, [_vp, _i, _vp, _vp, _vp]),
    "b200_clients_read_pre_dc": (_i, [_vp, _vp]),
    "b200_stream_prime": (_i, [_vp, _vp]),
    "b200_submit_block": (_i, [_vp, _pp, _i, _u64, _vp, _vp, _vp, _vp]),
    "b200_wait_block": (_i, [_vp]),
    "b200_waterfall_gather": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "b200_launch_count": (_u64, [_vp]),
    "b200_clients_fetch_async": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "b200_clients_fetch_wait": (_i, [_vp, _i]),
    "b200_set_waterfall_cadence": (_i, [_vp, _i]),
    "b200_set_frame_number": (_i, [_vp, _u64]),
    "b200_quant_table": (_i, [_i, _vp, _vp, _vp]),
    "b200_debug_tail_profile": (_i, [_vp, _i, _vp]),
}

_lib = None


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m phantomsdr_b200.build` "
                "(__graft_entry__.build()). The CUDA engine has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))

"""GPU: the C++ host driver (phantomsdr_b200/host/spectrum_loop.cpp - the reference's fft_task /
signal_loop / waterfall_loop call order over the C ABI, raw samples on stdin) against the golden
fixtures. Exercises the header-only adapter include/b200_fft.hpp exactly as a reference build would."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_input import CASES, raw_hop, client_table
from phantomsdr_b200 import FM, sizes
from phantomsdr_b200.build import HOST_BIN

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", list(CASES))
def test_cpp_driver_matches_golden(gpu_required, name):
    assert HOST_BIN.exists(), "run __graft_entry__.build()"
    case, g = CASES[name], np.load(GOLD / f"{name}.npz")
    cfg = case["cfg"]
    nframes = case["frames"]
    raw = b"".join(raw_hop(case, f).tobytes() for f in range(nframes + 2))
    specs = client_table(case)
    fmt = {"uint8": "u8", "int16": "s16"}[case["dtype"]]
    cmd = [str(HOST_BIN), "--sps", str(cfg.sps), "--fft", str(cfg.fft_size), "--format", fmt, "--frames", str(nframes)]
    if cfg.is_real:
        cmd.append("--real")
    for (l, mid, r, mode) in specs:
        cmd += ["--client", f"{l},{mid!r},{r},{mode}"]
    out = subprocess.run(cmd, input=raw, capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    buf, pos = out.stdout, 0
    h = cfg.audio_fft_size // 2
    audio, rows = {}, {}
    while pos < len(buf):
        magic, frame, a, b = struct.unpack_from("<4I", buf, pos)
        pos += 16
        if magic == 0x41554449:
            pwr = struct.unpack_from("<f", buf, pos)[0]
            pos += 4
            audio[(frame, a)] = (pwr, np.frombuffer(buf, np.int32, b, pos))
            pos += 4 * b
        else:
            assert magic == 0x57465241
            rows[frame] = (a, np.frombuffer(buf, np.int8, b, pos))
            pos += b
    assert len(audio) == nframes * len(specs)
    for (frame, i), (pwr, pcm) in audio.items():
        assert abs(pwr - g["pwr"][frame, i]) <= 1e-5 * g["pwr"][frame, i]
        if specs[i][3] != FM:
            assert np.abs(pcm - g["pcm"][frame, i]).max() <= 2
    # waterfall rows: sent every skip_num-th frame, from the last level (default client of websocket.cpp:195-198)
    R, L = cfg.fft_result_size, cfg.downsample_levels
    assert sorted(rows) == [f for f in range(nframes) if f % cfg.skip_num == 0]
    top = sizes.level_offset(L - 1, R) - sizes.level_offset(3, R)
    for frame, (level, row) in rows.items():
        assert level == L - 1
        want = g["q_hi"][frame][top: top + row.size]
        assert np.abs(row.astype(np.int32) - want.astype(np.int32)).max() <= 1

"""SURVEY 8d sweep: FFT size 2^16 .. 2^23 (c2c; sample rate scaled with the size so that the frame rate and the 360-point
audio FFT stay those of cfg 2) x clients {1, 64, 1024, 8192}: one device-timed bench.py leg each (no e2e, no CPU baseline).
Usage (on a GPU box): python tools/sweep.py > profiles/sweep_r2.json"""
import json
import subprocess
import sys

rows = []
for log2 in range(16, 24):
    for clients in (1, 64, 1024, 8192):
        sps = 35_000_000 * (1 << log2) // (1 << 20)
        cmd = [sys.executable, "bench.py", "--fft-log2", str(log2), "--sps", str(sps), "--clients", str(clients), "--steps", "20",
               "--warmup", "3", "--no-cpu-baseline", "--no-e2e"]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
            rows.append({"fft_log2": log2, "sps": sps, "clients": clients, "value_msps": d["value"], "ms_per_step": d["ms_per_step"],
                         "realtime_margin": d.get("realtime_margin"), "roofline_frac": d["roofline"]["frac"],
                         "forward_us_per_frame": d["roofline"]["us_per_frame"], "breakdown": d.get("breakdown"),
                         "audio_fft_size": d["config"]["audio_fft_size"], "clocks": d.get("clocks")})
        except Exception as exc:  # keep going: a failed cell is a finding, not a reason to lose the sweep
            rows.append({"fft_log2": log2, "sps": sps, "clients": clients, "error": repr(exc)[:300],
                         "stderr": (out.stderr[-400:] if "out" in dir() else "")})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
print(json.dumps({"what": "bench.py legs, device-timed, inputs resident in HBM, 64 frames per launch group", "rows": rows}, indent=1))

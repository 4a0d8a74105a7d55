"""Measured FP32 FMA throughput of this GPU for scalar FFMA and packed FFMA2 (vmp_fma_probe).
    python tools/fp32_peak.py  -> JSON line {ffma_tflops, ffma2_tflops, nominal_tflops, sm_mhz}"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import _lib  # noqa: E402


def measure(packed, grid=148 * 8, iters=4096, reps=5):
    lib = _lib.load()
    out = torch.empty(grid * 256, dtype=torch.float32, device='cuda')
    st = _lib.stream_ptr()
    lib.vmp_fma_probe(packed, grid, 64, _lib.ptr(out), st)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.vmp_fma_probe(packed, grid, iters, _lib.ptr(out), st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        flops = 2.0 * grid * 256 * iters * 128
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


if __name__ == '__main__':
    r = {'ffma_tflops': measure(0), 'ffma2_tflops': measure(1),
         'nominal_tflops': 148 * 128 * 2 * 1.965e9 / 1e12, 'gpu': torch.cuda.get_device_name(0)}
    print(json.dumps(r))

#!/usr/bin/env python
"""Where the CTA-pair kernels (tc_gemm2.cuh) spend their clocks - development probe, NOT a bench.

    TVAE_PROBE=1 python target-vae_b200/csrc/build.py        # -> tvae_b200/libtvae_b200_probe.so (clock64 probes compiled in)
    python tools/probe_pair.py cfg4 [B]

Runs 3 fwd+bwd steps of the config with the probe build and prints, per pair kernel, the share of the MMA issuer's
lifetime spent waiting for a drained accumulator (= epilogue serialised with the MMAs) and for operand stages (= starved
by the generator warps / TMA), the epilogue thread's wait / work split and the generator thread's wait / work split."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "target-vae_b200", "tvae_b200", "libtvae_b200_probe.so")
os.environ["TVAE_LIB"] = PROBE
for p in (ROOT, os.path.join(ROOT, "target-vae_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from tvae_b200 import _lib  # noqa: E402
from tvae_b200.config import PRESETS  # noqa: E402

NAMES = {1: "conv1_fwd", 2: "conv1_wgrad", 3: "gen_l1_fwd", 4: "gen_l1_wgrad", 5: "linear_tn"}


def main():
    cfg = PRESETS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
    ctx = bench.Ctx()
    wl = bench.Workload(ctx, cfg, B)
    lib = _lib.lib()
    for kv in sys.argv[3:]:                     # development knobs: id=value
        k, v = kv.split("=")
        lib.tvae_test_set_knob(int(k), int(v))
        print("knob", k, "=", v)
    buf = (ctypes.c_ulonglong * 128)()
    for i in range(2):
        wl.step_resident(i)
    _lib.check(lib.tvae_probe_read(buf), "tvae_probe_read")        # clears the counters
    K = 3
    for i in range(K):
        wl.step_resident(i)
    _lib.check(lib.tvae_probe_read(buf), "tvae_probe_read")
    print(f"{cfg.name} B={B}: clocks per launch, mean over the CTA pairs of a launch ({K} steps)")
    print(f"{'kernel':14s} {'mma life':>10s} {'wait acc':>9s} {'wait ops':>9s} {'issue':>7s} | {'epi wait':>9s} {'epi work':>9s} | {'gen wait':>9s} {'gen work':>9s}")
    for slot, name in NAMES.items():
        c = [buf[slot * 16 + j] for j in range(16)]
        if c[7] == 0:
            continue
        n = c[7]
        life = c[0] / n
        f = lambda v: f"{100.0 * v / n / life:8.1f}%"
        print(f"{name:14s} {life:10.0f} {f(c[1])} {f(c[2])} {f(c[0] - c[1] - c[2])[1:]} | {f(c[3])} {f(c[4])} | {f(c[5])} {f(c[6])}")
        if sum(c[8:14]):
            tot = float(sum(c[8:14]))
            seg = ["tmem ld wait", "convert+sts", "proxy fence + arrive", "wait for a free buffer", "-", "-"]
            print("      epilogue segments: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in zip(seg, c[8:14])))
        if c[14] or c[15]:
            print(f"      generator group 0: tile begin {f(c[14])}, per-chunk prepare {f(c[15])} (both include the wait for the other group at the slab-refill barrier)")
        if slot in (1, 3):
            r = [buf[(6 if slot == 1 else 7) * 16 + j] for j in range(16)]
            if r[3]:
                print(f"      store issuer, clocks per 64-column block: waits for the block {r[0] / r[3]:.0f}, issue + commit {r[1] / r[3]:.0f}, "
                      f"wait_read {r[2] / r[3]:.0f}   ({r[3] / n:.0f} blocks per launch and CTA)")
    print("(gen columns: generator group 0 of 2 = every other chunk; mma / epilogue columns: the leader CTA)")


if __name__ == "__main__":
    main()

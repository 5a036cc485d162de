#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Captures the host CPU's RCPPS look-up table by executing `_mm_rcp_ss` through the
compiled reference shim (oracle/_ref/liblpcnet_ref_A.so: ref_rcp_table/ref_rcp) and checks the structural claim
the engine relies on: rcp(2^e * 1.m) = T[m >> 12] * 2^-e, i.e. the result depends only on
the top 11 mantissa bits and the exponent is handled separately.

Writes tests/golden/rcpps_table.bin (2048 little-endian u32) + rcpps_table.json (provenance: CPU model).
`vec_avx.h:406,435` (`_mm256_rcp_ps` inside tanh8_approx / sigmoid8_approx) is the only consumer on the hot path.
"""
import ctypes, json, os, sys
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "_ref", "liblpcnet_ref_A.so"))
lib.ref_rcp.restype = ctypes.c_float
lib.ref_rcp.argtypes = [ctypes.c_float]
tab = np.zeros(2048, dtype=np.uint32)
lib.ref_rcp_table(tab.ctypes.data_as(ctypes.c_void_p))

def emu(xbits):
    k = (xbits >> 12) & 0x7FF
    e = (xbits & 0x7F800000) - 0x3F800000
    return (tab[k].astype(np.int64) - e).astype(np.uint32)

rng = np.random.default_rng(7)
# exhaustive over mantissas at exponent 0, random over a wide exponent range
xs = np.concatenate([(0x3F800000 + np.arange(1 << 23, dtype=np.int64))[::97],
                     rng.integers(0x30000000, 0x4F000000, size=400000, dtype=np.int64)]).astype(np.uint32)
bad = 0
for chunk in np.array_split(xs, 50):
    got = np.array([np.float32(lib.ref_rcp(float(v))).view(np.uint32) for v in chunk.view(np.float32)], dtype=np.uint32)
    bad += int((got != emu(chunk.astype(np.int64))).sum())
cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
print("cpu:", cpu, "| checked", xs.size, "inputs, mismatches:", bad)
print("T[0]=%08x T[1]=%08x T[2047]=%08x" % (tab[0], tab[1], tab[2047]), "low 12 bits all zero:", bool(((tab & 0xFFF) == 0).all()))
if bad == 0:
    out = os.path.join(here, "..", "tests", "golden")
    tab.astype("<u4").tofile(os.path.join(out, "rcpps_table.bin"))
    json.dump({"cpu": cpu, "entries": 2048, "rule": "rcp(2^e*1.m) bits = T[m>>12] - (e<<23)",
               "T0": "%08x" % tab[0], "T2047": "%08x" % tab[2047]}, open(os.path.join(out, "rcpps_table.json"), "w"), indent=1)
    print("wrote tests/golden/rcpps_table.bin")
sys.exit(1 if bad else 0)

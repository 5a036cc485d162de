"""Shared test plumbing: builds/loads the checkers (oracle port, compiled reference) through ctypes.

The oracle and the compiled reference are CHECKERS: they are only ever used from tests/, smoke() and the
bench's cpu_baseline leg — never by the product library.
"""
import ctypes
import functools
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
GEN = os.path.join(ORACLE, "_gen")
GOLDEN = os.path.join(ROOT, "tests", "golden")
LPC_GAMMA = 0.9

c_p = ctypes.c_void_p


@functools.lru_cache(None)
def gen_dir():
    """Deterministically (re)generate the synthetic model files."""
    need = ["model_int8.bin", "model_float.bin", "codebooks.bin"]
    if not all(os.path.exists(os.path.join(GEN, n)) for n in need):
        import gen_model
        gen_model.generate(GEN, c_sources=os.path.isdir("/root/reference/src"))
    return GEN


@functools.lru_cache(None)
def blob(kind="int8"):
    return open(os.path.join(gen_dir(), "model_%s.bin" % kind), "rb").read()


@functools.lru_cache(None)
def codebooks():
    return np.fromfile(os.path.join(gen_dir(), "codebooks.bin"), dtype=np.float32)


@functools.lru_cache(None)
def rcp_table():
    return np.fromfile(os.path.join(GOLDEN, "rcpps_table.bin"), dtype="<u4")


@functools.lru_cache(None)
def oracle_lib():
    so = os.path.join(ORACLE, "_build", "liblpcnet_oracle.so")
    src = os.path.join(ORACLE, "lpcnet_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE, "port"], stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(so)
    L.oracle_model_create.restype = c_p
    L.oracle_model_create.argtypes = [ctypes.c_char_p, ctypes.c_int, c_p, ctypes.c_float, c_p]
    L.oracle_state_create.restype = c_p
    L.oracle_state_create.argtypes = [c_p]
    L.oracle_state_destroy.argtypes = [c_p]
    L.oracle_reset.argtypes = [c_p]
    L.oracle_synthesize.argtypes = [c_p, c_p, c_p, ctypes.c_int]
    L.oracle_synthesize_trace.argtypes = [c_p, c_p, c_p, ctypes.c_int, c_p]
    L.oracle_decode.argtypes = [c_p, c_p, c_p]
    L.oracle_decode_packet.argtypes = [c_p, c_p, c_p]
    L.oracle_frame_network.argtypes = [c_p, c_p, c_p, c_p, c_p]
    L.oracle_get_tables.argtypes = [c_p] * 6
    L.oracle_get_state.argtypes = [c_p] * 6
    L.oracle_tanh.restype = ctypes.c_float
    L.oracle_tanh.argtypes = [c_p, ctypes.c_float]
    L.oracle_sigmoid.restype = ctypes.c_float
    L.oracle_sigmoid.argtypes = [c_p, ctypes.c_float]
    L.oracle_lin2ulaw.argtypes = [ctypes.c_float]
    L.oracle_synthesize_batch.restype = ctypes.c_double
    L.oracle_synthesize_batch.argtypes = [c_p, c_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_p]
    L.oracle_decode_batch.restype = ctypes.c_double
    L.oracle_decode_batch.argtypes = [c_p, c_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_p]
    return L


@functools.lru_cache(None)
def oracle_model(kind="int8"):
    L = oracle_lib()
    b = blob(kind)
    m = L.oracle_model_create(b, len(b), rcp_table().ctypes.data, LPC_GAMMA, codebooks().ctypes.data)
    assert m, "oracle failed to parse the blob"
    return m


def oracle_synth(features, kind="int8", nthreads=8):
    """features [n][T][20] float32 -> pcm [n][T*160] int16 via the CPU restatement."""
    L = oracle_lib()
    f = np.ascontiguousarray(features, dtype=np.float32)
    n, T, stride = f.shape
    pcm = np.zeros((n, T * 160), dtype=np.int16)
    L.oracle_synthesize_batch(oracle_model(kind), f.ctypes.data, stride, n, T, nthreads, pcm.ctypes.data)
    return pcm


def oracle_decode(packets, kind="int8", nthreads=8):
    """packets [n][P][8] uint8 -> pcm [n][P*640] int16."""
    L = oracle_lib()
    p = np.ascontiguousarray(packets, dtype=np.uint8)
    n, P, _ = p.shape
    pcm = np.zeros((n, P * 640), dtype=np.int16)
    L.oracle_decode_batch(oracle_model(kind), p.ctypes.data, n, P, nthreads, pcm.ctypes.data)
    return pcm


def have_ref():
    return os.path.exists(os.path.join(ORACLE, "_ref", "liblpcnet_ref_A.so"))


@functools.lru_cache(None)
def ref_lib(build="A"):
    """The UNTOUCHED reference compiled by oracle/Makefile (`make ref`)."""
    L = ctypes.CDLL(os.path.join(ORACLE, "_ref", "liblpcnet_ref_%s.so" % build))
    L.ref_synth_stream.argtypes = [ctypes.c_char_p, ctypes.c_int, c_p, ctypes.c_int, ctypes.c_int, c_p]
    L.ref_decode_stream.argtypes = [ctypes.c_char_p, ctypes.c_int, c_p, ctypes.c_int, c_p]
    L.ref_decode_packet.argtypes = [c_p, c_p, c_p]
    L.ref_frame_network.argtypes = [ctypes.c_char_p, ctypes.c_int, c_p, ctypes.c_int, ctypes.c_int, c_p, c_p, c_p]
    L.ref_rcp.restype = ctypes.c_float
    L.ref_rcp.argtypes = [ctypes.c_float]
    L.ref_ulaw2lin.restype = ctypes.c_float
    L.ref_ulaw2lin.argtypes = [ctypes.c_float]
    L.ref_lin2ulaw.argtypes = [ctypes.c_float]
    L.ref_activation.argtypes = [c_p, c_p, ctypes.c_int, ctypes.c_int]
    L.ref_time_synthesis.restype = ctypes.c_double
    L.ref_time_synthesis.argtypes = [ctypes.c_char_p, ctypes.c_int, c_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_p]
    return L


def ref_synth(features, build="A"):
    L = ref_lib(build)
    b = blob("float" if build == "B" else "int8")
    f = np.ascontiguousarray(features, dtype=np.float32)
    n, T, stride = f.shape
    pcm = np.zeros((n, T * 160), dtype=np.int16)
    for s in range(n):
        assert L.ref_synth_stream(b, len(b), f[s].ctypes.data, stride, T, pcm[s].ctypes.data) == 0
    return pcm


def ref_decode(packets, build="A"):
    L = ref_lib(build)
    b = blob("float" if build == "B" else "int8")
    p = np.ascontiguousarray(packets, dtype=np.uint8)
    n, P, _ = p.shape
    pcm = np.zeros((n, P * 640), dtype=np.int16)
    for s in range(n):
        assert L.ref_decode_stream(b, len(b), p[s].ctypes.data, P, pcm[s].ctypes.data) == 0
    return pcm

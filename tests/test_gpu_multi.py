"""GPU tests (-m gpu) of multi-GPU inside the C-ABI: the one-process / N-device API (lpcnet_b200_multi_*), the PCM
sink that gathers a shard's output into a buffer owned by someone else (another device via peer access, another process via
CUDA IPC), all through the C-ABI, no torch.  Sharding must not change a single sample: stream s of the job is bit-identical
to stream s of a single-device batch (and hence to the oracle, tests/test_gpu_parity.py).  The >= 2 device cases skip on a
one-GPU box (`gpurun --gpus 2` runs them)."""
import ctypes
import os
import subprocess
import sys
import numpy as np
import pytest
import helpers as H
from fixtures import make_feature_batch, make_packets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import lpcnet_b200
    from lpcnet_b200 import build
    build.build()
    assert lpcnet_b200.device_count() > 0, "GPU test selected but no CUDA device is visible"
    return lpcnet_b200


def _single(eng, f, kind="int8"):
    b = eng.Batch(f.shape[0], H.blob(kind), lpc_gamma=H.LPC_GAMMA)
    out = b.synthesize(f)
    b.close()
    return out


def test_multi_on_one_device_equals_a_plain_batch_and_the_oracle(eng):
    n, T = 37, 20                                   # 20 frames: two chunks (16 + 4) through the sink path
    f = make_feature_batch(range(300, 300 + n), T)
    m = eng.Multi(n, H.blob("int8"), [0], config=(H.LPC_GAMMA, -1, -1))
    assert m.shard(0) == (0, 0, n)
    a = m.synthesize(f[:, :9])
    b = m.synthesize(f[:, 9:], gather=True)         # state carries over between the two call styles
    m.close()
    got = np.concatenate([a, b], axis=1)
    np.testing.assert_array_equal(got, _single(eng, f))
    np.testing.assert_array_equal(got[:5], H.oracle_synth(f[:5], "int8"))


def test_pcm_sink_places_rows_at_an_offset(eng):
    """lpcnet_b200_batch_set_pcm_sink: every chunk of a call also lands at sink[(first_row + s) * pitch + t]; rows outside the
    shard are untouched; removing the sink stops the forwarding."""
    L = eng.lib()
    n, T, rows, first = 6, 19, 11, 3
    pitch = T * 160 + 32
    f = make_feature_batch(range(40, 40 + n), T)
    sink = L.lpcnet_b200_device_alloc(rows * pitch * 2)
    fill = np.full((rows, pitch), 0x5555, np.int16)
    L.lpcnet_b200_memcpy_h2d(sink, fill.ctypes.data, fill.nbytes)
    b = eng.Batch(n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA)
    b.set_pcm_sink(sink, pitch, first)
    direct = b.synthesize(f)
    got = np.zeros_like(fill)
    L.lpcnet_b200_memcpy_d2h(got.ctypes.data, sink, got.nbytes)
    np.testing.assert_array_equal(got[first:first + n, :T * 160], direct)
    assert (got[:first] == 0x5555).all() and (got[first + n:] == 0x5555).all() and (got[:, T * 160:] == 0x5555).all()
    # too small a pitch is refused, NULL removes the sink
    b.set_pcm_sink(sink, 100, 0)
    with pytest.raises(eng.LPCNetB200Error, match="pitch"):
        b.synthesize(f)
    b.set_pcm_sink(None, 0, 0)
    L.lpcnet_b200_memcpy_h2d(sink, fill.ctypes.data, fill.nbytes)
    b.synthesize(f[:, :3])
    L.lpcnet_b200_memcpy_d2h(got.ctypes.data, sink, got.nbytes)
    assert (got == 0x5555).all()
    b.close()
    L.lpcnet_b200_device_free(sink)


_CHILD = r"""
import sys, numpy as np
sys.path[:0] = [%(root)r, %(root)r + "/tests", %(root)r + "/oracle"]
import helpers as H, lpcnet_b200
from fixtures import make_feature_batch
L = lpcnet_b200.lib()
handle = bytes.fromhex(sys.argv[1]); first, n, T, total, dev = map(int, sys.argv[2:7])
L.lpcnet_b200_set_device(dev)
p = L.lpcnet_b200_ipc_open(handle)
assert p, L.lpcnet_b200_last_error()
b = lpcnet_b200.Batch(n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA, device=dev)
b.set_pcm_sink(p, T * 160, first)
b.synthesize(make_feature_batch(range(700 + first, 700 + first + n), T))
b.set_pcm_sink(None, 0, 0)
b.close()
assert L.lpcnet_b200_ipc_close(p) == 0
print("child ok")
"""


def _gather_through_ipc(eng, child_dev):
    """One process per shard (the torchrun / MPI deployment): the owner exports its gather buffer, the other process opens it
    and sets it as its PCM sink."""
    L = eng.lib()
    total, T = 10, 17
    own_n, first = 4, 4                                   # the child fills rows 4..9, the owner rows 0..3
    f = make_feature_batch(range(700, 700 + total), T)
    L.lpcnet_b200_set_device(0)
    buf = L.lpcnet_b200_device_alloc_on(0, total * T * 160 * 2)
    zero = np.zeros((total, T * 160), np.int16)
    L.lpcnet_b200_memcpy_h2d(buf, zero.ctypes.data, zero.nbytes)
    handle = (ctypes.c_ubyte * 64)()
    assert L.lpcnet_b200_ipc_export(buf, handle) == 0
    r = subprocess.run([sys.executable, "-c", _CHILD % {"root": H.ROOT}, bytes(handle).hex(), str(first), str(total - own_n), str(T), str(total), str(child_dev)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stdout + r.stderr
    b = eng.Batch(own_n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA, device=0)
    b.set_pcm_sink(buf, T * 160, 0)
    b.synthesize(f[:own_n])
    b.close()
    got = np.zeros_like(zero)
    L.lpcnet_b200_memcpy_d2h(got.ctypes.data, buf, got.nbytes)
    L.lpcnet_b200_device_free(buf)
    np.testing.assert_array_equal(got, _single(eng, f))


def test_gather_across_processes_through_cuda_ipc_same_device(eng):
    _gather_through_ipc(eng, 0)


def test_multi_argument_checks(eng):
    with pytest.raises(eng.LPCNetB200Error, match="listed twice" if eng.device_count() >= 2 else "devices requested"):
        eng.Multi(8, H.blob("int8"), [0, 0])
    with pytest.raises(eng.LPCNetB200Error, match="devices requested|out of range"):
        eng.Multi(8, H.blob("int8"), list(range(eng.device_count() + 1)))
    with pytest.raises(eng.LPCNetB200Error, match="cannot be sharded"):
        eng.Multi(0, H.blob("int8"), [0])


# ---------------------------------------------------------------- >= 2 devices
def _need2(eng):
    if eng.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")


@pytest.mark.parametrize("n", [11, 64])
def test_two_devices_equal_one(eng, n):
    _need2(eng)
    T = 21
    f = make_feature_batch(range(500, 500 + n), T)
    want = _single(eng, f)
    m = eng.Multi(n, H.blob("int8"), [0, 1], config=(H.LPC_GAMMA, -1, -1))
    assert m.peer_access()
    lo = [m.shard(k) for k in range(2)]
    assert lo[0][1] == 0 and lo[1][1] == lo[0][2] and lo[0][2] + lo[1][2] == n and [d for d, _, _ in lo] == [0, 1]
    np.testing.assert_array_equal(m.synthesize(f[:, :10]), want[:, :1600])
    np.testing.assert_array_equal(m.synthesize(f[:, 10:], gather=True), want[:, 1600:])     # PCM gathered on device 0 over NVLink
    m.reset()
    np.testing.assert_array_equal(m.synthesize(f, gather=True), want)
    m.close()


def test_two_devices_decode_and_reversed_device_order(eng):
    _need2(eng)
    n, P = 9, 3
    pk = np.stack([make_packets(s, P) for s in range(n)])
    b = eng.Batch(n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA, codebooks=H.codebooks())
    want = b.decode(pk)
    b.close()
    m = eng.Multi(n, H.blob("int8"), [1, 0], config=(H.LPC_GAMMA, -1, -1), codebooks=H.codebooks())     # gather buffer on device 1
    np.testing.assert_array_equal(m.decode(pk), want)
    m.reset()
    np.testing.assert_array_equal(m.decode(pk, gather=True), want)
    m.close()


def test_gather_across_processes_through_cuda_ipc_other_device(eng):
    _need2(eng)
    _gather_through_ipc(eng, 1)

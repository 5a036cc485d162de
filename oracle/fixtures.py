"""Deterministic synthetic inputs (TEST INFRASTRUCTURE): feature frames and codec packets.

SURVEY.md 8(d) "Synthetic inputs": per-stream seed 1000+s for features, 2000+s for packets.  A feature frame
is the 20 floats `lpcnet_synthesize` consumes (include/lpcnet.h:188): 18 Bark cepstra, pitch, pitch-corr
(src/lpcnet.c:93 derives the pitch index from features[18]; src/freq.c:310 the LPC from features[0..17]).
"""
import numpy as np

NB_FEATURES = 20


def make_features(stream, nframes, key_every=8):
    """[nframes][20] float32, smooth random walk between key frames (speech-like ranges)."""
    rng = np.random.default_rng(1000 + int(stream))
    nkeys = nframes // key_every + 2
    keys = np.zeros((nkeys, NB_FEATURES))
    keys[:, 0] = rng.uniform(4.0, 11.0, nkeys)                   # c0 (log energy)
    decay = np.exp(-np.arange(1, 18) / 8.0)
    keys[:, 1:18] = rng.normal(0.0, 0.9, (nkeys, 17)) * decay    # c1..c17
    keys[:, 18] = rng.uniform(-1.3, 3.0, nkeys)                  # pitch feature: period = 50*f+100 in [35,250]
    keys[:, 19] = rng.uniform(-0.5, 0.5, nkeys)                  # pitch correlation - 0.5
    t = np.arange(nframes) / key_every
    i0 = np.floor(t).astype(int)
    fr = (t - i0)[:, None]
    feat = (1 - fr) * keys[i0] + fr * keys[i0 + 1]
    feat[:, :18] += rng.normal(0.0, 0.03, (nframes, 18))
    return np.ascontiguousarray(feat, dtype=np.float32)


def make_feature_batch(streams, nframes):
    """[n_streams][nframes][20] float32 for stream ids `streams`."""
    return np.stack([make_features(s, nframes) for s in streams])


def make_packets(stream, npackets):
    """[npackets][8] uint8 uniform random codec packets (src/lpcnet_dec.c:81 consumes 64 bits each)."""
    rng = np.random.default_rng(2000 + int(stream))
    return rng.integers(0, 256, size=(npackets, 8), dtype=np.uint8)


def make_pcm(stream, nframes):
    """[nframes*160] int16 speech-like test signal for the analysis side: harmonic source with vibrato
    (f0 80..300 Hz), a couple of formant-ish resonances, amplitude modulation with silent gaps, plus noise; seed 3000+s."""
    rng = np.random.default_rng(3000 + int(stream))
    n = nframes * 160
    t = np.arange(n) / 16000.0
    f0 = rng.uniform(80, 300) * (1 + 0.05 * np.sin(2 * np.pi * rng.uniform(3, 7) * t + rng.uniform(0, 6.28)))
    f0 *= np.exp(0.15 * np.cumsum(rng.normal(0, 0.002, n)))
    ph = 2 * np.pi * np.cumsum(f0) / 16000.0
    src = sum(np.sin(k * ph + rng.uniform(0, 6.28)) / k for k in range(1, 24))
    # two resonances (direct-form recursion on the excitation)
    y = src + 0.3 * rng.normal(0, 1, n)
    for fc, bw in ((rng.uniform(400, 900), 120.0), (rng.uniform(1400, 2600), 200.0)):
        r = np.exp(-np.pi * bw / 16000.0)
        a1, a2 = 2 * r * np.cos(2 * np.pi * fc / 16000.0), -r * r
        out = np.zeros(n)
        y1 = y2 = 0.0
        for i in range(n):
            v = y[i] + a1 * y1 + a2 * y2
            out[i] = v
            y2, y1 = y1, v
        y = y + 0.5 * out
    env = np.clip(0.55 + 0.6 * np.sin(2 * np.pi * rng.uniform(0.7, 2.5) * t + rng.uniform(0, 6.28)), 0, 1) ** 2   # gaps of near-silence
    y = y / (np.abs(y).max() + 1e-9) * rng.uniform(3000, 15000) * env + rng.normal(0, 12, n)
    return np.clip(np.rint(y), -32767, 32767).astype(np.int16)


def make_pcm_batch(streams, nframes):
    return np.stack([make_pcm(s, nframes) for s in streams])

#!/usr/bin/env python3
"""Golden vectors of the analysis side, produced by the UNTOUCHED reference (oracle/_ref build A: src/lpcnet_enc.c
compiled by oracle/Makefile) on the deterministic synthetic PCM of oracle/fixtures.py (seed 3000+s) and the synthetic VQ codebooks.
Run in the build container only.  The input PCM is stored with the outputs so that the fixture does not depend on numpy's libm.

  enc_A.npz : pcm[8][40*160] int16
              features[8][40][36]   lpcnet_compute_single_frame_features per frame      (`lpcnet_demo -features`)
              packets[8][10][8]     lpcnet_encode per 640 samples                        (`lpcnet_demo -encode`)
              features4[8][40][36]  lpcnet_compute_features per 640 samples (unquantised superframe analysis)
              mixed_packets[2][3][8], mixed_features[2][28][36]  3 x lpcnet_encode then 28 x single-frame analysis on ONE state
              half_window[160], dct_table[324]  the reference's tables (src/lpcnet_tables.c)
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import helpers as H
from fixtures import make_pcm_batch

pcm = make_pcm_batch(range(8), 40)
L = H.ref_lib("A")
mp = np.zeros((2, 3, 8), np.uint8); mf = np.zeros((2, 28, 36), np.float32)
for s in range(2):
    L.ref_encode_then_features(pcm[s].ctypes.data, 3, mp[s].ctypes.data, 28, mf[s].ctypes.data)
hw = np.zeros(160, np.float32); dct = np.zeros(324, np.float32)
L.ref_enc_tables(hw.ctypes.data, dct.ctypes.data)
out = dict(pcm=pcm, features=H.ref_features(pcm), packets=H.ref_encode(pcm), features4=H.ref_features4(pcm),
           mixed_packets=mp, mixed_features=mf, half_window=hw, dct_table=dct)
np.savez_compressed(os.path.join(HERE, "enc_A.npz"), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)

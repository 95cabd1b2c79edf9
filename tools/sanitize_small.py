"""Small encode + decode through the C ABI for compute-sanitizer (tools/sanitize.sh): 2 streams x 320x192 x 4
pictures (IDR + P: every stage of k_encode_mbs, the deblocking / expansion / pack kernels), pipelined with two
batches in flight, then the GPU decoder on the produced streams; checks the decoder reproduces the encoder's
reconstruction.  Exit code 0 = functional result correct (the sanitizer reports its own findings)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import h264lib
from openh264_b200.binding import BatchEncoder, BatchDecoder
W, H, N, S = 320, 192, 4, 2
clips = [h264lib.synth_clip(W, H, N, seed=40 + s) for s in range(S)]
fsz = W * H * 3 // 2
enc = BatchEncoder(W, H, qp=28, fps=30.0, n_streams=S)
aus = [[] for _ in range(S)]
enc.submit([c[:fsz] for c in clips])
for f in range(1, N + 1):
    if f < N:
        enc.submit([c[f * fsz:(f + 1) * fsz] for c in clips])
    bs, _ = enc.collect()
    for s in range(S):
        aus[s].append(bytes(bs[s]))
rec = [enc.recon(s) for s in range(S)]
enc.close()
dec = BatchDecoder(W, H, n_streams=S)
for f in range(N):
    pics = dec.decode([aus[s][f] for s in range(S)])
dec.close()
for s in range(S):
    assert np.array_equal(pics[s], rec[s]), "decoder output != encoder reconstruction"
print("sanitize_small ok: %d bytes of bitstream" % sum(len(a) for au in aus for a in au))

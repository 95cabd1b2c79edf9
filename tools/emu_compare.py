"""Debug driver: encode a clip with the host emulation build and with the reference; compare per frame."""
import ctypes as C, hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from h264_dump import split_nals

def load():
    E = C.CDLL(os.path.join(ROOT, "tests/emu/libb2h264_emu.so"))
    E.emu_encode.restype = C.c_long
    E.emu_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    R = C.CDLL(os.path.join(ROOT, "oracle/_ref/librefshim.so"))
    R.ref_encode.restype = C.c_long
    R.ref_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_long, C.c_void_p, C.POINTER(C.c_double)]
    R.ref_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    return E, R

def run(yuv, w, h, n, qp=26, fps=12.0, complexity=2):
    E, R = load()
    cap = 64 << 20
    o1, o2 = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    f1, f2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    rec = np.zeros(n * w * h * 3 // 2, np.uint8)
    secs = C.c_double()
    t2 = R.ref_encode(yuv.ctypes.data, w, h, n, qp, complexity, 1, fps, o2.ctypes.data, cap, f2.ctypes.data, C.byref(secs))
    t1 = E.emu_encode(yuv.ctypes.data, w, h, n, qp, fps, o1.ctypes.data, cap, f1.ctypes.data, rec.ctypes.data)
    return o1[:t1], f1, o2[:t2], f2, rec, R

if __name__ == "__main__":
    clip = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/res/CiscoVT2people_320x192_12fps.yuv"
    w, h = int(sys.argv[2]) if len(sys.argv) > 2 else 320, int(sys.argv[3]) if len(sys.argv) > 3 else 192
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    yuv = np.fromfile(clip, dtype=np.uint8)
    o1, f1, o2, f2, rec, R = run(yuv, w, h, n)
    print("emu frame bytes", f1.tolist())
    print("ref frame bytes", f2.tolist())
    same = len(o1) == len(o2) and bool((o1 == o2).all())
    print("IDENTICAL" if same else "DIFFERENT")
    if not same:
        m = min(len(o1), len(o2))
        d = np.nonzero(o1[:m] != o2[:m])[0]
        print("first diff at byte", int(d[0]) if len(d) else m, "of", len(o1), len(o2))
        # decode the reference stream and compare recon of frame(s) to locate the first differing MB
        dec = np.zeros(n * w * h * 3 // 2 + 64, np.uint8); W = C.c_int(); H = C.c_int(); s = C.c_double()
        nf = R.ref_decode(o2.ctypes.data, len(o2), dec.ctypes.data, len(dec), C.byref(W), C.byref(H), C.byref(s))
        fsz = w * h * 3 // 2
        for fi in range(min(n, nf)):
            a = rec[fi * fsz: fi * fsz + w * h].reshape(h, w); b = dec[fi * fsz: fi * fsz + w * h].reshape(h, w)
            if not (a == b).all():
                ys, xs = np.nonzero(a != b)
                mbs = sorted(set((int(y) // 16) * ((w + 15) // 16) + int(x) // 16 for y, x in zip(ys, xs)))
                print(f"frame {fi}: luma recon differs in {len(mbs)} MBs, first MB idx {mbs[0]} (x={mbs[0] % ((w+15)//16)}, y={mbs[0] // ((w+15)//16)})")
                break
        else:
            print("recon luma equal for all compared frames")

def bits_of(nal, nbytes=24):
    from h264_dump import unescape
    d = unescape(bytes(nal))
    return ''.join(f'{b:08b}' for b in d[1:1 + nbytes])

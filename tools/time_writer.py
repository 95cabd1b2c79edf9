"""Microbenchmark of the HOST entropy coders (CAVLC / CABAC slice writers + NAL encapsulation) on one 1080p P picture of the bench's
"hard" content, through the host emulation build: microseconds per access unit on one core.   python tools/time_writer.py"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import h264lib
E = C.CDLL(os.path.join(ROOT, "tests", "emu", "libb2h264_emu.so"))
E.emu_encode.restype = C.c_long
E.emu_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
E.emu_last_writer_us.restype = C.c_double
E.emu_last_parser_us.restype = C.c_double
w, h, n = 1920, 1080, 2
yuv = h264lib.synth_clip(w, h, n, seed=5, noise=int(sys.argv[1]) if len(sys.argv) > 1 else 8)
for cabac in (0, 1):
    E.emu_set_entropy(cabac, 0); E.emu_set_time_writer(5)
    out, fb = np.zeros(64 << 20, np.uint8), np.zeros(n, np.int32)
    tot = E.emu_encode(yuv.ctypes.data, w, h, n, 26, 30.0, out.ctypes.data, out.size, fb.ctypes.data, None)
    print("cabac" if cabac else "cavlc", "P picture bytes", fb[1], "writer us/AU %.0f" % E.emu_last_writer_us(), "-> %.0f Mbit/s per core" % (fb[1] * 8 / E.emu_last_writer_us()),
          "| parser us/AU %.0f -> %.0f Mbit/s" % (E.emu_last_parser_us(), fb[1] * 8 / E.emu_last_parser_us()))
E.emu_set_time_writer(0); E.emu_set_entropy(0, 66)

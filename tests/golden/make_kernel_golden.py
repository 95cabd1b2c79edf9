"""Writes tests/golden/kernels.json: SHA-1 of every seeded kernel case (tests/kernel_cases.py)
as computed by the UNMODIFIED reference through oracle/_ref/librefshim.so.  Run in the container
that has /root/reference after `make -f oracle/Makefile.ref`:

    python tests/golden/make_kernel_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import h264lib  # noqa: E402
from kernel_cases import CASES, run_case  # noqa: E402

assert h264lib.have_ref(), "build oracle/_ref first (make -f oracle/Makefile.ref)"
out = {name: h264lib.sha1(*run_case(name, h264lib.ref())) for name in sorted(CASES)}
with open(os.path.join(HERE, "kernels.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out, indent=1))

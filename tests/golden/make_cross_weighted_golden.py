"""Writes tests/golden/cross_weighted.json: SHA-1s of the oracle's outputs on the seeded cases of tests/test_cross_weighted.py.
Run only after tests/test_cross_weighted.py::test_oracle_*_matches_reference passed on a machine with oracle/_ref (the oracle is
then equal to the unmodified reference on exactly these cases)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_cross_weighted as t

json.dump(t.oracle_digests(), open(os.path.join(HERE, "cross_weighted.json"), "w"), indent=1, sort_keys=True)
print("written")

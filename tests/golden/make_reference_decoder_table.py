"""Extracts the reference's own decoder golden table — (bitstream, SHA-1 of the decoded I420) pairs of
test/api/decoder_test.cpp:90-142 — into tests/golden/reference_decoder_hashes.json.  The hashes are what the
reference's DecoderOutputTest expects; tests/test_reference_build.py checks that the build in oracle/_ref
reproduces them, i.e. that the compiled oracle IS the reference."""
import json, os, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/api/decoder_test.cpp"
pairs = re.findall(r'\{"(res/[^"]+)",\s*"([0-9a-f]{40})"\}', open(src).read())
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_decoder_hashes.json")
json.dump({"source": "test/api/decoder_test.cpp (kFileParamArray)", "pairs": pairs}, open(out, "w"), indent=1)
print(len(pairs), "pairs ->", out)

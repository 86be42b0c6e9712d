#!/usr/bin/env python
"""Golden digests for the text sink: sha256 and length of what the UNMODIFIED reference
sink (oracle/_ref/libwtref.so: TeeWiggleIterator + runWiggleIterator,
/root/reference/src/wigWriter.c) prints for the seeded segment lists of
tests/test_writer.py.  Run here (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_writer_golden.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_writer import CASES, NAMES, make_segments  # noqa: E402

ref = Oracle("reference")
out = {}
for n, seed in CASES:
    c, s, e, v = make_segments(n, seed)
    for bg in (0, 1):
        text = ref.write(c, s, e, v, NAMES, bool(bg))
        out["%d/%d/%d" % (n, seed, bg)] = {"sha256": hashlib.sha256(text).hexdigest(), "bytes": len(text)}
with open(os.path.join(ROOT, "tests", "golden", "writer_digests.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", len(out), "digests")

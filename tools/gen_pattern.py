"""Regenerates orb_slam3_fast_b200/csrc/orb_pattern.inc (the 256x4 rBRIEF test-pair table).

The table is a numeric constant of the ORB algorithm (Rublee et al. 2011; OpenCV orb.cpp `bit_pattern_31_`;
reference src/ORBextractor.cc:149-406). Run only in the build container where /root/reference is mounted.
"""
import re, sys
src = open(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/ORBextractor.cc").read()
s = src.index("bit_pattern_31_[256 * 4] = {"); e = src.index("};", s)
body = re.sub(r"/\*.*?\*/", "", src[s:e].split("{", 1)[1], flags=re.S)
v = [int(x) for x in re.findall(r"-?\d+", body)]
assert len(v) == 1024
for i in range(0, 1024, 16):
    print("  " + " ".join("%d,%d,%d,%d," % tuple(v[j:j + 4]) for j in range(i, i + 16, 4)))

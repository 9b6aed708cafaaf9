#!/usr/bin/env python
"""Build isonclust2_b200/data/p_emp_prob.bin from the reference's empirical table.

The table P(minimizer shared | k, w, e1, e2) is input DATA the mapped-fraction
test cannot run without (reference: src/p_emp_prob_data.h, 41 880 text rows
`k w p e1 e2`, consumed by InitMinSharedMap, src/p_emp_prob.cpp:22-47).  This
script parses the rows (float() is correctly rounded, like the istream >>
double the reference uses) and stores them, IN FILE ORDER (later rows
overwrite earlier ones in the reference), as a flat little-endian record array:

    magic 'ICBPEMP1', u32 n_rows, then n_rows x {u8 k, u8 w, u8 e1c, u8 e2c, f64 p}

with e?c = round(e*100) (1..15).  It asserts that every e parses to exactly
c/100.0, which is what makes "round to 2 decimals then std::map lookup"
(p_emp_prob.cpp:66-94) equal to an index lookup.

Run in the build container (needs /root/reference); the output is committed.
"""
import re
import struct
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/p_emp_prob_data.h"
out = sys.argv[2] if len(sys.argv) > 2 else "isonclust2_b200/data/p_emp_prob.bin"
rows = []
for line in open(src):
    m = re.match(r'\s*"(\d+)\t(\d+)\t(\S+)\t(\S+)\t(\S+)\\n"', line)
    if not m:
        continue
    k, w, p, e1, e2 = int(m[1]), int(m[2]), float(m[3]), float(m[4]), float(m[5])
    c1, c2 = round(e1 * 100), round(e2 * 100)
    assert e1 == c1 / 100.0 and e2 == c2 / 100.0 and 1 <= c1 <= 15 and 1 <= c2 <= 15, line
    rows.append((k, w, c1, c2, p))
with open(out, "wb") as f:
    f.write(b"ICBPEMP1")
    f.write(struct.pack("<I", len(rows)))
    for r in rows:
        f.write(struct.pack("<BBBBd", *r))
print(len(rows), "rows ->", out)

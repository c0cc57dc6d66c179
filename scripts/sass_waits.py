#!/usr/bin/env python
"""Lists, for one kernel of the library, the instructions inside a SASS address range whose wait mask includes a
scoreboard that global loads in the same range write -- i.e. where a prefetch can be waited for by accident.
usage: sass_waits.py <mangled-name-substring> [lo_hex hi_hex]"""
import re
import subprocess
import sys

lib = "tractor3d_b200/csrc/libt3d_particles.so"
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = [m for m in re.findall(r"Function : (\S+)", names) if sys.argv[1] in m][0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout.split("\n")
ins = []
i = 0
while i < len(sass) - 1:
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", sass[i])
    if m:
        hi = int(re.search(r"/\* (0x[0-9a-f]+) \*/", sass[i + 1]).group(1), 16)
        c = hi >> 41
        ins.append((int(m.group(1), 16), m.group(2).strip(), (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3F))
        i += 2
    else:
        i += 1
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi_ = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
print(fn[:80], len(ins), "instructions")
for a, t, wr, rd, wait in ins:
    if lo <= a <= hi_ and (wait or wr != 7 and ("LDG" in t or "LD.E" in t)):
        print(f"{a:05x} wr={wr if wr != 7 else '-'} wait={wait:06b}  {t[:90]}")

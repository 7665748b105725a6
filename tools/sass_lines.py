#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel in an object file (nvdisasm line info; needs -lineinfo).
usage: sass_lines.py object.o kernel_mangled_substring [top_n]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

obj, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
m = re.search(r"\.text\.(\S*%s\S*):" % re.escape(kname), dis)
sec = dis[dis.index(".text.%s:" % m.group(1)):]
nxt = sec.find("//--------------------- .text.", 10)
sec = sec[: nxt if nxt > 0 else len(sec)]
cur, counts, ops = ("?", 0), collections.Counter(), collections.defaultdict(collections.Counter)
for ln in sec.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    mi = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if mi:
        counts[cur] += 1
        ops[cur][mi.group(1).split(".")[0]] += 1
print("total", sum(counts.values()))
for (f, l), c in sorted(counts.items(), key=lambda kv: (kv[0][0], kv[0][1]))[:10000]:
    print(f"{f}:{l:4d} {c:4d}  " + " ".join(f"{k}x{v}" for k, v in ops[(f, l)].most_common(8)))

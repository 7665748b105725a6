#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: per source line warp-instructions, thread-instructions, SIMT efficiency.
usage: ncu_lines.py report.ncu-rep object.o kernel_mangled_substring [top_n]"""
import csv, io, re, subprocess, sys, os, tempfile, glob

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# section of the kernel
m = re.search(r"\.text\.(\S*%s\S*):" % re.escape(kname), dis)
name = m.group(1)
sec = dis[dis.index(".text.%s:" % name):]
nxt = sec.find("//--------------------- .text.", 10)
sec = sec[: nxt if nxt > 0 else len(sec)]
line_of = []
cur = ("?", 0)
for ln in sec.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        line_of.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + os.environ.get("NCU_KERNEL", ".*"), "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
for k, r in enumerate(data):  # several launches of the kernel in the report: the first one
    if r and r[0] == "Kernel Name":
        data = data[:k]
        break
assert len(data) == len(line_of), (len(data), len(line_of))
agg = {}
for r, lo in zip(data, line_of):
    a = agg.setdefault(lo, [0, 0, 0])
    a[0] += int(r[ci["Instructions Executed"]])
    a[1] += int(r[ci["Thread Instructions Executed"]])
    a[2] += int(r[ci["# Samples"]])
tot_i = sum(a[0] for a in agg.values()); tot_t = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_i:.3e} thread-inst {tot_t:.3e} eff {tot_t/tot_i:.2f}/32 samples {tot_s}")
srcs = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if f not in srcs:
        p = [x for x in glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), f))]
        srcs[f] = open(p[0]).read().splitlines() if p else []
    text = srcs[f][l - 1].strip()[:90] if srcs[f] and l - 1 < len(srcs[f]) else ""
    print(f"{f}:{l:4d} inst {100*a[0]/tot_i:5.1f}%  eff {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/max(tot_s,1):5.1f}%  | {text}")

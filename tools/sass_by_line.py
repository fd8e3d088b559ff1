#!/usr/bin/env python
"""tools/sass_by_line.py OBJECT.o MANGLED_KERNEL_SUBSTRING COUNTS.csv.gz : joins the per-instruction executed counts of an ncu capture
(tools/ncu_capture.sh -> *_sass.csv.gz) with the line info of the same kernel in the object file (nvdisasm -g) and prints warp
instructions executed per source line, heaviest first."""
import csv, gzip, re, subprocess, sys, collections, tempfile, os, glob
obj, kern, counts = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + '/*.cubin')[0]
txt = subprocess.run(['nvdisasm', '-c', '-g', cubin], capture_output=True, text=True).stdout.split('\n')
start = [n for n, l in enumerate(txt) if l.startswith('.text.') and kern in l]
if not start:
    sys.exit("kernel not found")
lines = []       # (source line of each instruction, text)
cur = '?'
for l in txt[start[0] + 1:]:
    if l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if 'inlined' not in m.group(3):
            pass
        cur = os.path.basename(m.group(1)) + ':' + m.group(2)
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append((cur, l.strip()))
rows = list(csv.reader(gzip.open(counts, 'rt')))
print("instructions in object:", len(lines), " in capture:", len(rows))
agg = collections.Counter(); tot = 0
for (src, _), r in zip(lines, rows):
    n = int(r[2] or 0); agg[src] += n; tot += n
for k, v in agg.most_common(45):
    print("%-28s %12d %5.1f%%" % (k, v, 100.0 * v / tot))

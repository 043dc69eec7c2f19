"""Development aid: the branch regions (BSSY) and calls of one kernel with the source lines they come from, and the
sizes of the basic blocks in between (no GPU needed).  usage: sass_branches.py <kernel-name-substring> [lib.so]"""
import os
import re
import subprocess
import sys
import tempfile

pat = sys.argv[1]
lib = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else "rsplash_b200/libsplash_cuda.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
secs = re.split(r"\n//-+ \.text\.", txt)
for sec in secs[1:]:
    name = sec.split(" ", 1)[0]
    if pat not in name:
        continue
    print("==", name[-70:])
    last, n, since = None, 0, 0
    for l in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            last = (m.group(1).split("/")[-1], m.group(2))
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            n += 1
            since += 1
            if re.search(r"\b(BSSY|CALL|BRA|BSYNC|RET|EXIT)\b", l):
                ins = re.sub(r"\s+", " ", l.split("*/", 1)[1]).strip()[:58]
                print(f"{n:6d} +{since:4d}  {ins:60s} {last}")
                since = 0

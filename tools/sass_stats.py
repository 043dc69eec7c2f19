"""Development aid: static SASS statistics of the library's kernels (no GPU needed).
usage: sass_stats.py [lib.so] [kernel-name-substring ...]
Prints, per matching kernel (and the device functions it calls), the instruction count and the opcode mix."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "rsplash_b200/libsplash_cuda.so"
pats = sys.argv[2:] or ["k_splash_fusedIfLb1ELb1", "k_run_bulkIfLb1"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if not any(p in name for p in pats):
        continue
    ops = collections.Counter()
    n = 0
    for line in f.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s", line)
        if m:
            op = m.group(2)
            full = op + (m.group(3) or "")
            if op == "IMAD" and ".MOV" in full:
                op = "IMAD.MOV"
            ops[op] += 1
            n += 1
    print(f"== {name[:110]}: {n} instructions ({n * 16 / 1024:.1f} KB)")
    print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))

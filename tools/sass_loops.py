"""Development aid: list the loops (backward branches) of one kernel in an object file with their instruction mix.
usage: sass_loops.py <object> <kernel-name-substring> [--dump LO HI]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ins = None, []
for l in names.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(t):
    p = t.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
print(len(ins), "instructions")
for a, t in ins:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        body = [x for x in ins if lo <= x[0] <= a]
        h = collections.Counter(op(x[1]) for x in body)
        print("loop %#x..%#x: %d instr; %s" % (lo, a, len(body), " ".join("%s=%d" % kv for kv in h.most_common(14))))
if "--dump" in sys.argv:
    i = sys.argv.index("--dump"); lo, hi = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
    for a, t in ins:
        if lo <= a <= hi: print("%#06x  %s" % (a, t))

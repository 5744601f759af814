"""Per-instruction view of an `ncu --page source --csv --print-source sass` dump: the hottest instructions by stall samples,
and per-stall-reason totals.  usage: ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in body)
texec = sum(f(r, "Instructions Executed") for r in body)
print("total samples", tot, "total warp instr", texec)
reasons = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print({k: round(100 * sum(f(r, k) for r in body) / tot, 1) for k in reasons if sum(f(r, k) for r in body) / tot > 0.005})
order = sorted(range(len(body)), key=lambda i: -f(body[i], "# Samples"))[:top]
for i in sorted(order):
    r = body[i]
    rs = sorted(((f(r, k), k) for k in reasons), reverse=True)[:2]
    print("%5d %5.2f%% exec %9d  %-60s %s" % (i, 100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"), r[ix["Source"]].strip()[:60],
                                       " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(f(r, "# Samples"), 1)) for v, k in rs)))

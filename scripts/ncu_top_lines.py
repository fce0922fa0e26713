"""stdin: `ncu -i rep --page source --csv`; prints the source lines with the most warp-stall samples"""
import csv, sys
rows = list(csv.reader(sys.stdin))
if not rows:
    sys.exit(0)
# the page starts with "Kernel Name", ... lines; the table header is the first row that has a "Source" column
hi = next((i for i, r in enumerate(rows) if any(x.strip() == "Source" for x in r)), 0)
h = rows[hi]
rows = rows[hi:]
def col(name):
    for i, x in enumerate(h):
        if x.strip() == name:
            return i
    return None
src = col("Source"); samp = col("# Samples") or col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)")
if samp is None:
    samp = next((i for i, x in enumerate(h) if "Sampl" in x), None)
inst = col("Instructions Executed")
if src is None or samp is None:
    print("columns:", h[:30]); sys.exit(0)
out = []
for r in rows[1:]:
    try:
        out.append((int(float(r[samp].replace(",", "") or 0)), r[src].strip()[:140], r[inst] if inst is not None else ""))
    except Exception:
        pass
tot = sum(x[0] for x in out) or 1
for n, s, i in sorted(out, reverse=True)[:25]:
    print("%6.2f %%  %8d  %s" % (100.0 * n / tot, n, s))

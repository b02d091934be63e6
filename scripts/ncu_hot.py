"""Top stalled instructions of an exported `ncu --page source --csv --print-source sass` file (.csv or .csv.gz)."""
import csv, gzip, io, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def num(r, n):
    try: return float(r[col[n]])
    except ValueError: return 0.0
total = sum(num(r, "# Samples") for r in body)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
print("total samples", total, "instructions", len(body), "executed", sum(num(r, "Instructions Executed") for r in body))
agg = {n: sum(num(r, n) for r in body) for n in stalls}
print({k: round(v / max(total, 1), 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((num(r, n), n) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {num(r,'# Samples')/max(total,1)*100:5.1f}%  exec {int(num(r,'Instructions Executed')):9d}  {r[col['Source']][:90]:<90} {why[0][1]}:{int(why[0][0])} {why[1][1]}:{int(why[1][0])}")

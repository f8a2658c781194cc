"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data = []
for r in rows:
    if not r or r[0] in ("File Path", "Function Name", "Line No") or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    idx = None
    for i in range(1, len(r) - 1):
        if r[i] == "-" and r[i + 1] == "-":
            idx = i
            break
    if idx is None:
        continue
    src = ",".join(r[1:idx])
    vals = r[idx + 2:]
    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    samples, notissued, inst = f(vals[0]), f(vals[1]), f(vals[3])
    data.append((inst, samples, ln, src))
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("total warp-inst %.4g  stall samples %.4g" % (tot, tots))
for inst, s, ln, src in sorted(data, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  L%4d  %s" % (100 * inst / tot, 100 * s / max(tots, 1), ln, src.strip()[:120]))

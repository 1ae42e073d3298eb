"""Per-source-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv`.
usage: ncu_lines.py X.csv [min_percent] [function-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
only = sys.argv[3] if len(sys.argv) > 3 else ""
cur = fn = None
data = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": fn = r[1].split("(")[0]; continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try: data.setdefault(fn, []).append((cur, int(r[0]), r[1].strip(), int(r[7]), int(r[6])))
        except ValueError: pass
for fn, d in data.items():
    if only not in fn: continue
    tot = sum(x[3] for x in d); ts = sum(x[4] for x in d)
    print(f"=== {fn}: warp instructions {tot}, samples {ts}")
    for f, ln, src, ins, smp in d:
        if 100 * ins / tot >= thr or 100 * smp / max(ts, 1) >= thr:
            print(f"{100*ins/tot:5.1f}% inst {100*smp/max(ts,1):5.1f}% smp | {f}:{ln} | {src[:130]}")

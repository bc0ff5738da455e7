"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, out, hdr = "", [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, sm, ti = r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
    elif hdr and len(r) > ii and r[0].isdigit() and r[2] == "-":   # a CUDA source line (SASS rows carry an address)
        n = int(r[ii]) if r[ii].isdigit() else 0
        if n:
            out.append((n, int(r[sm]) if r[sm].isdigit() else 0, int(r[ti]) if r[ti].isdigit() else 0, fname, r[0], r[1].strip()[:110]))
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp-instructions {tot}, samples {tots}")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]:>10d} {100*o[0]/tot:5.1f}%  smp {100*o[1]/max(tots,1):5.1f}%  thr/inst {o[2]/o[0]:5.1f}  {o[3]}:{o[4]:>4}  {o[5]}")

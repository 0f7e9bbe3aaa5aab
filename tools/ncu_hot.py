"""Top SASS instructions by stall samples from an .ncu-rep: python tools/ncu_hot.py file.ncu-rep [kernel-substr] [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for n, hi in enumerate(his):
    name = " ".join(rows[hi - 1][:2])
    if want not in name:
        continue
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    print(name, "samples", tot)
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:N]
    for i in sorted(order):
        r = body[i]
        print(f"{i:5d} {int(r[ix['# Samples']] or 0):7d} {100*int(r[ix['# Samples']] or 0)/max(1,tot):5.1f}%  exec {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']][:110]}")
    break

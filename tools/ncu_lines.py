"""Per-source-line stall samples of one kernel in an .ncu-rep (compiled with -lineinfo): top lines by samples.
usage: ncu_lines.py rep [kernel-id like :::2] [top]"""
import csv, subprocess, sys
kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, H, acc = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        H = r
    elif H and len(r) == len(H) and r[0].strip().isdigit():
        i_s, i_i = H.index("# Samples"), H.index("Instructions Executed")
        try:
            acc.append((int(r[i_s]), int(r[i_i]), fname, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(a[0] for a in acc) or 1
print("total samples", tot)
for s, ins, f, ln, src in sorted(acc, reverse=True)[:top]:
    print(f"{100 * s / tot:5.1f}% inst={ins:8d} {f}:{ln}  {src}")

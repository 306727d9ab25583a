"""Per-phase (barrier-delimited) executed-instruction and stall-sample shares of the last kernel in an .ncu-rep (source page)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        ks.append((r[1], cur))
        continue
    if cur is not None:
        cur.append(r)
name, k = ks[-1]
hdr, body = k[0], k[1:]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
allI, allS = sum(int(r[iI]) for r in body), sum(int(r[iSm]) for r in body)
print(name[:60], "sass", len(body), "instr", allI, "samples", allS)
split = sys.argv[2] if len(sys.argv) > 2 else "BAR.SYNC"
ph = acc_i = acc_s = n = 0
top = []
for r in body:
    acc_i += int(r[iI]); acc_s += int(r[iSm]); n += 1
    top.append((int(r[iSm]), int(r[iI]), r[iS].strip()))
    if split in r[iS]:
        print(f"phase {ph}: sass={n} instr={acc_i} ({100*acc_i/allI:.1f}%) samples={acc_s} ({100*acc_s/allS:.1f}%)")
        ph += 1; acc_i = acc_s = n = 0
print(f"phase {ph}: sass={n} instr={acc_i} ({100*acc_i/allI:.1f}%) samples={acc_s} ({100*acc_s/allS:.1f}%)")
if len(sys.argv) > 3:
    for s, i, src in sorted(top, reverse=True)[:int(sys.argv[3])]:
        print(f"{s:6d} {i:9d}  {src}")

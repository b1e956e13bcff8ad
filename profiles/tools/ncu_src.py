"""Per-instruction executed counts from an ncu report's source page, grouped into address ranges."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[2:], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    n = int(r[iex]); tot += n
    lines.append((a - base, r[isrc].strip(), n, int(r[ismp] or 0)))
print("total warp-instructions", tot)
for off, s, n, smp in lines:
    print(f"{off:05x} {n:>12d} {smp:>6d}  {s}")

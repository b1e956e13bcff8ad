"""Extract one function's SASS from the built library: python scratch/sass_fn.py <substr> [start end]"""
import os, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "video-stab_b200", "libvstab_b200.so")], capture_output=True, text=True).stdout
i = txt.index("Function : " + [m for m in re.findall(r"Function : (\S+)", txt) if sys.argv[1] in m][0])
j = txt.find("Function :", i + 10)
body = txt[i:j if j > 0 else None]
ins = []
for ln in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
if len(sys.argv) <= 2:
    print(len(ins), "instructions")
    for a, s in ins:
        if re.search(r"\bBAR|BRA|EXIT|STL|LDL|CALL|RET", s):
            print(f"{a:04x} {s}")
else:
    for a, s in ins:
        if lo <= a < hi:
            print(f"{a:04x} {s}")

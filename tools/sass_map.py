"""Layout of a kernel's SASS by 4 KB block (development): python tools/sass_map.py LIB 'kernel-substring'"""
import re, collections, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; rows = []
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            t = m.group(2).split(); op = t[1] if t[0].startswith("@") else t[0]
            rows.append((int(m.group(1), 16), op, m.group(2)))
print(len(rows), "instructions,", len(rows) * 16 // 1024, "KB")
blk = collections.defaultdict(collections.Counter)
for a, op, _ in rows:
    c = blk[a // 4096]; c["n"] += 1
    if re.match(r"D(MUL|ADD|FMA|SETP)", op): c["f64"] += 1
    if re.match(r"F(FMA|MUL|ADD|MNMX|SEL|SETP)", op): c["f32"] += 1
    if "LDG" in op: c["ldg"] += 1
    if re.match(r"(VOTE|REDUX|SHFL|WARPSYNC|ATOMS|ATOMG|RED|MATCH)", op): c["coll"] += 1
    if op.startswith("BRA"): c["bra"] += 1
    if op[:3] in ("LDS", "STS"): c["smem"] += 1
    if op[:3] in ("LDL", "STL"): c["local"] += 1
    if op.startswith("CALL"): c["call"] += 1
for b in sorted(blk):
    c = blk[b]
    print("%3dKB: f32 %3d f64 %3d ldg %2d smem %2d coll %2d bra %2d local %2d call %d" % (b * 4, c["f32"], c["f64"], c["ldg"], c["smem"], c["coll"], c["bra"], c["local"], c["call"]))
mx = 0
for a, op, txt in rows:
    m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", txt)
    if m and op.startswith("BRA"):
        t = int(m.group(1), 16)
        if t < a and a - t > mx:
            mx = a - t; print("back edge %x -> %x (%.1f KB span)" % (a, t, (a - t) / 1024))

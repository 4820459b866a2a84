"""Static SASS size of a kernel's address range by source line (development): which statements of the per-warp engine
the instructions of the hot loop come from.  python tools/sass_lines.py LIB kernel-substring lo hi [caller-line]
(nvdisasm -gi prints the inlining chain in front of every instruction; with caller-line, instructions are attributed to
the statement of the function inlined at that line -- e.g. the line of runWarp's body -- instead of the innermost one)"""
import collections, os, re, subprocess, sys, tempfile
lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
caller = int(sys.argv[5]) if len(sys.argv) > 5 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cnt = collections.Counter(); total = 0
for f in os.listdir(tmp):
    if not f.startswith("se3_kernels"): continue
    out = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside = False; chain = []; fresh = False
    for l in out.splitlines():
        if l.startswith(".text."):
            inside = pat in l; chain = []; fresh = False; continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            if not fresh:      # (the chain is only printed when it changes: it holds until the next one)
                chain = []; fresh = True
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), int(m.group(4)) if m.group(4) else None)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m:
            fresh = False
            a = int(m.group(1), 16)
            if lo <= a <= hi and chain:
                key = chain[0][:2]
                if caller is not None:
                    for fl, ln, at in chain:
                        if at == caller: key = (fl, ln); break
                cnt[key] += 1; total += 1
    break
src = {}
print("%d instructions (%.1f KB) in [%x, %x]" % (total, total * 16 / 1024, lo, hi))
for (fl, ln), c in cnt.most_common(60):
    if fl not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mplambda_b200", "csrc", fl)
        src[fl] = open(p).read().split("\n") if os.path.exists(p) else []
    txt = src[fl][ln - 1].strip()[:95] if 0 < ln <= len(src[fl]) else ""
    print("%5d %5.1f%%  %-16s:%5d | %s" % (c, 100.0 * c / total, fl, ln, txt))

"""Static evidence per kernel from the in-tree objects: registers / spills / shared memory (cuobjdump -res-usage) and
SASS mnemonic counts (cuobjdump -sass).  No GPU involved.  Usage: python profiles/sass_table.py rNN"""
import collections, glob, os, re, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sorted(glob.glob(os.path.join(ROOT, "interpolatory_b200", "_lib", "*.o")))
MNEM = ("VABSDIFF", "VABSDIFF4", "UTMALDG", "LDGSTS", "SYNCS", "REDUX", "ATOMS", "ATOMG", "STL", "LDL")
def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"^void ", "", o).split("(")[0] for o in out]
rows = []
for o in objs:
    res = subprocess.run(["cuobjdump", "-res-usage", o], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    cur = None
    counts = collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if cur and m:
            op = m.group(1)
            counts[cur]["instrs"] += 1
            if op == "VABSDIFF4" or (op == "VABSDIFF" and (m.group(2) or "").startswith(".U8") is False and op == "VABSDIFF4"):
                counts[cur]["VABSDIFF4"] += 1
            elif op == "VABSDIFF":
                counts[cur]["VABSDIFF"] += 1
            elif op in MNEM:
                counts[cur][op] += 1
    names = list(usage)
    for mangled, nice in zip(names, demangle(names)):
        if nice.startswith("ub_kernel"):
            continue
        u, c = usage[mangled], counts[mangled]
        rows.append((os.path.basename(o).replace(".o", ".cu"), nice, u.get("REG", 0), u.get("STACK", 0), u.get("SHARED", 0), c))
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.md"), "w") as f:
    f.write(f"# Static evidence, {tag}: resource usage and SASS mnemonics per kernel (sm_100a)\n\n")
    f.write("Generated in the build container from the in-tree objects (`cuobjdump -res-usage`, `cuobjdump -sass`) by "
            "`python profiles/sass_table.py " + tag + "`; no GPU involved.  STACK > 0 = local-memory frame (spills or arrays), "
            "STL / LDL = the instructions that touch it; UTMALDG = TMA tile loads (cp.async.bulk.tensor), SYNCS = mbarrier operations.\n\n")
    f.write("| kernel | registers | stack bytes | static smem (incl. 1 KB the driver reserves) | VABSDIFF | VABSDIFF4 | UTMALDG (TMA) | LDGSTS | SYNCS (mbarrier) | REDUX | ATOMS | ATOMG | STL | LDL | instrs |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for src, nice, reg, stack, shared, c in sorted(rows, key=lambda r: (r[0], r[1])):
        f.write(f"| `{nice}` ({src}) | {reg} | {stack} | {shared} | {c['VABSDIFF']} | {c['VABSDIFF4']} | {c['UTMALDG']} | {c['LDGSTS']} | {c['SYNCS']} | {c['REDUX']} | {c['ATOMS']} | {c['ATOMG']} | {c['STL']} | {c['LDL']} | {c['instrs']} |\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.md")).read())

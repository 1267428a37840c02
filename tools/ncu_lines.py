"""Per-source-line stall samples of one kernel launch in an ncu report, joined with the line table of the cubin.
usage: ncu_lines.py report.ncu-rep <launch index> <library.so> [top N]
(ncu's CLI prints per-instruction samples for SASS only; nvdisasm -g gives SASS offset -> source line.)"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, kid, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = out.splitlines()
kname = next(csv.reader([lines[0]]))[1]
rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
base = int(rows[0]["Address"], 16)
mangled = None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# demangle-free match: find the section whose demangled name matches
syms = re.findall(r"\.section\s+\.text\.(\S+?),", dis)
dem = subprocess.run(["cu++filt"] + syms, capture_output=True, text=True).stdout.splitlines()
want = re.sub(r"\(int\)", "", kname).replace("void ", "").replace("mor::", "")
for s, d in zip(syms, dem):
    d2 = re.sub(r"\(int\)", "", d).replace("void ", "").replace("mor::", "")
    if d2.split("(")[0] == want.split("(")[0]:
        mangled = s
        break
assert mangled, (kname, dem[:5])
sec = dis[dis.index(f".section\t.text.{mangled},"):]
nxt = sec.find("//--------------------- .text.", 10)
sec = sec[:nxt] if nxt > 0 else sec
line_of = {}
cur = None
stack = []
for l in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = collections.Counter()
inst = collections.Counter()
stall = collections.defaultdict(collections.Counter)
tot = 0
for r in rows:
    if not (r["# Samples"] or "").isdigit() or not r["Address"].startswith("0x"):
        continue
    s = int(r["# Samples"])
    if (r.get("Instructions Executed") or "").isdigit():
        inst[line_of.get(int(r["Address"], 16) - base, ("?", 0))] += int(r["Instructions Executed"])
    if not s:
        continue
    off = int(r["Address"], 16) - base
    key = line_of.get(off, ("?", 0))
    agg[key] += s
    tot += s
    for k, v in r.items():
        if k.startswith("stall_") and v not in ("", "0"):
            stall[key][k[6:]] += int(v)
src_cache = {}
def src(f, n):
    for root in ("dynamicslamtool_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:105] if 0 < n <= len(src_cache[p]) else ""
    return ""
print(kname, "samples", tot)
for (f, n), s in agg.most_common(top):
    st = ", ".join(f"{k}:{v}" for k, v in stall[(f, n)].most_common(3))
    print(f"{100*s/tot:5.1f}%  {f}:{n:<5} {src(f, n):105s} [{st}]")
if "--inst" in sys.argv:
    ti = sum(inst.values()) or 1
    print("warp instructions executed", ti)
    for (f, n), c in inst.most_common(top):
        print(f"{100*c/ti:5.1f}%  {c:9d}  {f}:{n:<5} {src(f, n):105s}")

"""SASS instructions per source line of one kernel (where does the code size come from?).
usage: sass_lines.py library.so <substring of the mangled kernel name> [top N]"""
import collections, os, re, subprocess, sys, tempfile
lib, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
secs = re.split(r"(?=//-+ \.text\.)", dis)
sec = next(s for s in secs if s.startswith("//") and want in s.split("\n", 1)[0])
cnt = collections.Counter(); cur = None; total = 0
for l in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l) and cur: cnt[cur] += 1; total += 1
print(sec.split("\n", 1)[0], "instructions:", total, "=", total * 16, "bytes")
src = {}
for (f, n), c in cnt.most_common(top):
    p = os.path.join("dynamicslamtool_b200/csrc", f)
    if os.path.exists(p) and p not in src: src[p] = open(p).read().splitlines()
    text = src[p][n - 1].strip()[:110] if p in src and n <= len(src[p]) else ""
    print(f"{c:6d}  {f}:{n:<5} {text}")

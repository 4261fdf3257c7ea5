"""Static instruction count per source function of one kernel in an object file (nvdisasm -g line info).
usage: python tools/sass_lines.py <obj.o> <kernel-name-substring> <source.cuh>   -- counts SASS instructions per function of
source.cuh (function = nearest preceding line that looks like a definition) and per other file."""
import collections, re, subprocess, sys, os, tempfile

obj, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
defs = []
for n, l in enumerate(open(src), 1):
    m = re.match(r"(?:__host__ __device__|__device__|static|template).*?\b(\w+)\s*\(", l)
    if m and not l.startswith("template <"):
        defs.append((n, m.group(1)))
def func(line):
    name = "?"
    for n, f in defs:
        if n <= line: name = f
    return name
cnt = collections.Counter(); kinds = collections.defaultdict(collections.Counter)
insec = False; cur = None
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        insec = kern in l; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1)); cur = func(int(m.group(2))) if f == os.path.basename(src) else f; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        cnt[cur] += 1; kinds[cur][m.group(1)] += 1
tot = sum(cnt.values())
print("total", tot, "instructions =", tot * 16 // 1024, "KB")
for k, v in cnt.most_common():
    print("%6d %-22s %s" % (v, k, " ".join("%s:%d" % kv for kv in kinds[k].most_common(6))))

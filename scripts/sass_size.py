"""Static SASS size of a kernel by source function / line range (code-footprint budget; the optimizer kernel is
instruction-fetch bound when its hot loops outgrow the 32 KB L1.5 instruction cache).

    python scripts/sass_size.py <kernel-substring> [source-file-suffix] [bucket]
Needs the in-tree libalore_b200.so built with -lineinfo."""
import re, subprocess, sys, tempfile, os, glob
kern = sys.argv[1]
suffix = sys.argv[2] if len(sys.argv) > 2 else "traj_opt.cuh"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("ALORE_B200_LIB") or os.path.join(root, "alore_legged_manipulator_b200", "libalore_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if os.path.basename(c).startswith("traj_opt")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(root, "alore_legged_manipulator_b200", "csrc", suffix)).read().split("\n")
funcs = [(i, m.group(1)) for i, l in enumerate(src, 1)
         for m in [re.match(r"^(?:__device__|__global__|template|static).*?(\w+)\((?!.*;\s*$)", l)] if m]
def fn(line):
    name = "?"
    for i, n in funcs:
        if i <= line: name = n
    return name
insec, cur = False, None
cnt, total = {}, 0
for l in out:
    if l.startswith("//---") and ".text." in l:
        insec = kern in l
        cur = None
        continue
    if not insec: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        total += 1
        key = cur if cur else ("?", 0)
        cnt[key] = cnt.get(key, 0) + 1
print(f"kernel ~{kern}: {total} SASS instructions = {total * 16 / 1024:.1f} KB")
agg = {}
for (f, ln), c in cnt.items():
    k = (f, fn(ln)) if f == suffix else (f, "*")
    agg[k] = agg.get(k, 0) + c
for k, c in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {c:6d}  {k[0]}:{k[1]}")
if len(sys.argv) > 3:
    b = int(sys.argv[3]); bk = {}
    for (f, ln), c in cnt.items():
        if f == suffix: bk[ln // b * b] = bk.get(ln // b * b, 0) + c
    for k in sorted(bk): print(f"  lines {k:5d}+  {bk[k]}")

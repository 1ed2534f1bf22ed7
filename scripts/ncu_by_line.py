"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line / device function."""
import csv, re, sys
path, srcfile = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
data = {}   # (file, line) -> [samples, instr, text]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; isamp = hdr.index("# Samples"); iins = hdr.index("Instructions Executed"); continue
    if hdr is None or cur_file is None: continue
    if r[0] in ("-", ""): continue        # SASS rows under a line
    try: ln = int(r[0])
    except ValueError: continue
    d = data.setdefault((cur_file, ln), [0.0, 0.0, r[1]])
    def fl(x):
        try: return float(x)
        except ValueError: return 0.0
    d[0] += fl(r[isamp]); d[1] += fl(r[iins])
tot_s = sum(v[0] for v in data.values()); tot_i = sum(v[1] for v in data.values())
print(f"total samples {tot_s:.0f}  instructions {tot_i:.3g}")
src = open(srcfile).read().split("\n")
funcs = [(i, m.group(1)) for i, l in enumerate(src, 1) for m in [re.match(r"^(?:__device__|__global__|template).*?(\w+)\(", l)] if m]
def fn(line):
    name = "?"
    for i, n in funcs:
        if i <= line: name = n
    return name
agg = {}
for (f, ln), v in data.items():
    key = fn(ln) if f.endswith(srcfile.split("/")[-1]) else f.split("/")[-1]
    a = agg.setdefault(key, [0.0, 0.0]); a[0] += v[0]; a[1] += v[1]
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} samples {s/tot_s*100:6.2f}%   instr {i/tot_i*100:6.2f}%")
print()
for (f, ln), v in sorted(data.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"  {f.split('/')[-1]}:{ln:<5d} samp {v[0]/tot_s*100:5.2f}% instr {v[1]/tot_i*100:5.2f}%  {v[2].strip()[:100]}")

"""Warp-stall samples of one kernel aggregated by CUDA source line — developer tool.

    python tools/ncu_hot_lines.py <rep.ncu-rep> <lib.so> <mangled-kernel-substring> <out.txt> [top]

`ncu --page source --csv` lists SASS instructions with their samples but without line numbers; the
line table comes from `nvdisasm -g` of the same cubin (built with -lineinfo); rows are matched by index."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern, out = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
inst = [r for r in rows[2:] if len(r) == len(hdr)]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = []   # one cubin per translation unit: take the one that holds the kernel
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern in d:
            dis = d.split("\n")
            break
lines, cur, inside = [], ("?", 0), False
for ln in dis:
    if ln.startswith("//---------------------"):
        inside = (".text." in ln and kern in ln)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
n = min(len(lines), len(inst))
agg = collections.Counter()
reasons = collections.defaultdict(collections.Counter)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
for k in range(n):
    s = float(inst[k][ci["# Samples"]] or 0)
    agg[lines[k]] += s
    tot += s
    for h in stall_cols:
        v = float(inst[k][ci[h]] or 0)
        if v:
            reasons[lines[k]][h[6:]] += v
src_cache = {}
def src(f, l):
    if f not in src_cache:
        for root in ("polymer-stats_b200/csrc", "/usr/local/cuda/include", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src_cache[f] = open(p, errors="replace").read().split("\n")
                break
        else:
            src_cache[f] = []
    t = src_cache[f]
    return t[l - 1].strip()[:110] if 0 < l <= len(t) else ""
with open(out, "w") as f:
    f.write(f"{rows[0][1]}\nSASS instructions: ncu {len(inst)}, nvdisasm {len(lines)}; total samples {tot:.0f}\n")
    tr = collections.Counter()
    for c in reasons.values():
        tr.update(c)
    f.write("stall mix: " + ", ".join(f"{k} {100*v/sum(tr.values()):.1f}%" for k, v in tr.most_common(8)) + "\n\n")
    for (fl, l), s in agg.most_common(top):
        rs = ", ".join(f"{k} {100*v/max(1,sum(reasons[(fl,l)].values())):.0f}%" for k, v in reasons[(fl, l)].most_common(3))
        f.write(f"{100*s/tot:6.2f}%  {fl}:{l:<5d} [{rs}]  {src(fl, l)}\n")
print(open(out).read())

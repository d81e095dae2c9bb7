"""Executed warp instructions of one kernel by source region and pipe class (developer tool; cf. ncu_hot_lines.py).
    python tools/ncu_inst_regions.py <rep.ncu-rep> <lib.so> <mangled-kernel-substring> [units]
`units` divides the counts (e.g. chains × trials of the profiled launch) so that the table reads "per trial"."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
inst = [r for r in rows[2:] if len(r) == len(hdr)]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = []
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
assert len(lines) == len(inst), (len(lines), len(inst))
def cls(sass):
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return "fp64"
    if op == "MUFU": return "xu"
    if op in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "LDC", "LDSM", "ATOMS", "ATOMG", "RED"): return "mem"
    if op in ("SHFL", "VOTE", "MATCH", "BAR", "WARPSYNC", "BSYNC", "BSSY", "BRA", "EXIT", "CALL", "RET", "NANOSLEEP", "BREAK", "WARPSYNC"): return "ctl"
    return "int/other"
agg = collections.defaultdict(collections.Counter)
thr = collections.Counter()
for k, r in enumerate(inst):
    f, l = lines[k]
    key = f"{f}:{l // 10 * 10:4d}"
    n = float(r[ci["Instructions Executed"]] or 0)
    agg[key][cls(r[ci["Source"]].strip())] += n
    thr[key] += float(r[ci["Thread Instructions Executed"]] or 0)
tot = collections.Counter()
for c in agg.values():
    tot.update(c)
print(f"{rows[0][1]}\nwarp instructions executed / {units:g}:  " + "  ".join(f"{k} {v / units:.1f}" for k, v in tot.most_common()),
      f"  total {sum(tot.values()) / units:.1f}")
print(f"{'region (10 source lines)':34s} {'total':>9s} {'fp64':>9s} {'xu':>7s} {'mem':>8s} {'ctl':>8s} {'int':>8s} {'lanes':>6s}")
for key, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:45]:
    t = sum(c.values())
    print(f"{key:34s} {t / units:9.1f} {c['fp64'] / units:9.1f} {c['xu'] / units:7.1f} {c['mem'] / units:8.1f} {c['ctl'] / units:8.1f} "
          f"{c['int/other'] / units:8.1f} {thr[key] / max(t, 1):6.1f}")

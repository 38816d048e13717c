"""Stall samples of one kernel of an ncu report, aggregated by CUDA source line (debug aid).
Joins the SASS page of the report (instruction i at offset 16*i) with `nvdisasm -g` line annotations of the library's cubin.
usage: python scripts/ncu_lines.py report.ncu-rep <kernel substring> [launch index among matches] [top N]"""
import csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
heads = [i for i, x in enumerate(r) if x and x[0] == "Address" and pat in r[i - 1][1]]
hi = heads[which]
mangled_hint = r[hi - 1][1]
hdr = r[hi]
rows = []
for x in r[hi + 1:]:
    if not x or not x[0].startswith("0x"):
        break
    rows.append(x)
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
# line table from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "quake_b200/lib/libquake_b200.so")], cwd=tmp, capture_output=True)
lines = {}
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, table = None, None, None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            cur_fn = m.group(1); table = lines.setdefault(cur_fn, {}); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and table is not None:
            table[int(m.group(1), 16)] = cur_line
# pick the function whose instruction count matches and whose demangled name matches the pattern
cands = [fn for fn, t in lines.items() if pat.split("<")[0].split("::")[-1] in fn and len(t) == len(rows)]
if not cands:
    cands = [fn for fn, t in lines.items() if pat.split("<")[0].split("::")[-1] in fn]
fn = cands[min(which if len(cands) > 1 and False else 0, len(cands) - 1)]
table = lines[fn]
agg = {}
tot = sum(int(x[isamp]) for x in rows); totex = sum(int(x[iex]) for x in rows)
for i, x in enumerate(rows):
    key = table.get(16 * i)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(x[isamp]); a[1] += int(x[iex]); a[2] += 1
print(f"{mangled_hint[:80]}  [{fn[:60]}] samples {tot} executed {totex} instrs {len(rows)}")
src_cache = {}
def src(key):
    if not key: return ""
    f, l = key
    for d in ("quake_b200/csrc",):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if l - 1 < len(src_cache[p]) else ""
    return ""
for key, a in sorted(agg.items(), key=lambda t: -t[1][0])[:N]:
    print(f"{100 * a[0] / max(tot,1):5.1f}% samp {100 * a[1] / max(totex,1):5.1f}% exec {a[2]:4d} ins  {key[0] if key else '?'}:{key[1] if key else 0:<5d} {src(key)}")

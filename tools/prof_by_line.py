"""Dev tool: instruction counts and stall samples per CUDA source line (via nvdisasm -g line markers)."""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1]
kern = os.environ.get('KERNEL', 'rc_obs_kernel')
nenv = float(os.environ.get("NENV", "655360"))
lib = "dynenv_b200/libdynenv_b200.so"
os.system("rm -rf /tmp/cub && mkdir -p /tmp/cub && cd /tmp/cub && cuobjdump -xelf all %s > /dev/null 2>&1" % os.path.abspath(lib))
os.system("ncu -i %s --page source --csv 2>/dev/null > /tmp/src.csv" % rep)
rows = list(csv.reader(open('/tmp/src.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
col = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
out = subprocess.run("cd /tmp/cub && for f in *.cubin; do nvdisasm -g -c $f; done", shell=True, capture_output=True, text=True).stdout.splitlines()
lines, insec, cur = [], False, ("?", 0)
for ln in out:
    if ln.startswith('//---') and '.text.' in ln:
        insec = kern in ln
        continue
    if ln.startswith('//---'):
        insec = False
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
agg = collections.defaultdict(lambda: [0, 0])
base = int(data[0][col['Address']], 16)
for r in data:
    k = (int(r[col['Address']], 16) - base) // 16
    key = lines[k] if k < len(lines) else ("?", 0)
    agg[key][0] += int(r[col['Instructions Executed']] or 0)
    agg[key][1] += int(r[col['# Samples']] or 0)
src = {}
for f in set(k[0] for k in agg):
    for d in ("dynenv_b200/csrc", "include"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            src[f] = open(p).read().splitlines()
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print("total inst per unit: %.0f" % (tot / nenv))
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:int(os.environ.get("TOP", "45"))]:
    text = src.get(key[0], [""] * 100000)[key[1] - 1].strip()[:110] if key[1] else ""
    print("%7.1f %5.1f%%s  %s:%d  %s" % (a[0] / nenv, 100 * a[1] / max(tots, 1), key[0], key[1], text))

"""Dev tool (here): aggregates the stall samples of an ncu report (--page source) per device function of the
step kernel, using the function labels of the disassembled cubin."""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof.ncu-rep"
lib = sys.argv[2] if len(sys.argv) > 2 else "dynenv_b200/libdynenv_b200.so"
os.system("rm -rf /tmp/cub && mkdir -p /tmp/cub && cd /tmp/cub && cuobjdump -xelf all %s > /dev/null 2>&1" % os.path.abspath(lib))
os.system("ncu -i %s --page source --csv 2>/dev/null > /tmp/src.csv" % rep)
rows = list(csv.reader(open('/tmp/src.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
col = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
out = subprocess.run("cd /tmp/cub && for f in *.cubin; do nvdisasm -c $f; done", shell=True, capture_output=True, text=True).stdout.splitlines()
funcs, n, insec = [], 0, False
for ln in out:
    if ln.startswith('//---') and '.text.' in ln:
        insec = os.environ.get('KERNEL','rc_phys_kernel') in ln
        n = 0
        if insec:
            funcs.append((0, os.environ.get('KERNEL','rc_phys_kernel')))
        continue
    if ln.startswith('//---'):
        insec = False
    if not insec:
        continue
    m = re.match(r'^(\$[_A-Za-z0-9$]+):\s*$', ln)
    if m:
        funcs.append((n, m.group(1)))
    if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', ln):
        n += 1
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
base = int(data[0][col['Address']], 16)
fi = 0
for r in data:
    k = (int(r[col['Address']], 16) - base) // 16
    while fi + 1 < len(funcs) and funcs[fi + 1][0] <= k:
        fi += 1
    nm = re.sub(r'^\$_ZN\d+_GLOBAL[^$]*\$', '', funcs[fi][1])
    a = agg[nm]
    a[0] += int(r[col['# Samples']] or 0)
    a[1] += int(r[col['Instructions Executed']] or 0)
    a[2] += int(r[col['stall_no_inst']] or 0)
    a[3] += int(r[col['Thread Instructions Executed']] or 0)
    a[4] += int(r[col['stall_wait']] or 0) + int(r[col['stall_branch_resolving']] or 0)
    a[5] += int(r[col['stall_short_sb']] or 0) + int(r[col['stall_long_sb']] or 0)
tot = sum(a[0] for a in agg.values())
toti = sum(a[1] for a in agg.values())
nenv = float(os.environ.get("NENV", "65536"))
print("%-52s %8s %8s %12s %8s %8s %8s %6s" % ("function", "samples%", "inst%", "inst/envstep", "no_inst%", "wait+br%", "sb%", "lanes"))
for nm, a in sorted(agg.items(), key=lambda x: -x[1][0])[:26]:
    dn = subprocess.run(['c++filt', nm], capture_output=True, text=True).stdout.strip()
    dn = re.sub(r'dynenv::(rc::)?', '', dn)
    dn = (dn.split('$_ZN6raster')[-1] if '$_ZN6raster' in dn else dn)[:52]
    print("%-52s %8.1f %8.1f %12.0f %8.0f %8.0f %8.0f %6.1f" % (dn, 100 * a[0] / tot, 100 * a[1] / toti, a[1] / nenv, 100 * a[2] / max(a[0], 1),
                                                      100 * a[4] / max(a[0], 1), 100 * a[5] / max(a[0], 1), a[3] / max(a[1], 1)))
print('total warp-instructions per env-step: %.0f' % (toti / nenv))

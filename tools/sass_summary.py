"""SASS instruction counts per kernel (dev tool): cuobjdump -sass lib.so | python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import sys, re, collections, subprocess
cur = None
cnt = collections.defaultdict(collections.Counter)
order = []
KEYS = ('DMMA', 'DFMA', 'FFMA', 'UBLKCP', 'SYNCS', 'REDUX', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'UCGABAR', 'CCTL')
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        if cur not in cnt:
            order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    c = cnt[cur]
    c['total'] += 1
    for key in KEYS:
        if op.startswith(key):
            c[key] += 1
names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
print("# SASS instruction counts per kernel of libgaunegf_b200.so (cuobjdump -sass, sm_100a)")
print("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), UBLKCP = bulk async copy (cp.async.bulk, the TMA engine's 1-D path),")
print("# SYNCS = mbarrier operations, REDUX = warp reductions (pivot search), UCGABAR = cluster barrier, CCTL = prefetch / cache control")
keys = ('total',) + KEYS
print("%-64s " % "kernel" + " ".join("%7s" % k for k in keys))
tot = collections.Counter()
for f, n in zip(order, names):
    c = cnt[f]
    tot.update(c)
    print("%-64s " % re.sub(r"\(.*", "", n)[:64] + " ".join("%7d" % c[k] for k in keys))
print("%-64s " % ("ALL (%d kernels)" % len(order)) + " ".join("%7d" % tot[k] for k in keys))

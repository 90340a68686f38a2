"""SASS mnemonic counts per kernel: cuobjdump -sass pinn_depthestimation_b200/libpinn_b200.so | python tools/sass_excerpt.py"""
import collections
import re
import sys

cur = None
cnt = collections.defaultdict(collections.Counter)
pat = re.compile(r'\b(UTCHMMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|REDG[.\w]*|RED\.[.\w]*|SYNCS[.\w]*|UCGABAR[.\w]*|USETMAXREG[.\w]*|UTMALDG[.\w]*)')
for l in sys.stdin:
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for x in pat.findall(l):
            cnt[cur][x] += 1
print("SASS mnemonic counts per kernel (cuobjdump -sass libpinn_b200.so, r2c build): tcgen05 MMA = UTCHMMA, TMEM load = LDTM, "
      "1-D TMA bulk copy = UBLKCP, commit = UTCBAR, setmaxnreg = USETMAXREG; the issue loops are rolled (one UTCHMMA per MMA of a stage)")
for k, c in cnt.items():
    if any(x.startswith(('UTCHMMA', 'UBLKCP', 'LDTM')) for x in c):
        print(k)
        for x, n in sorted(c.items()):
            print(f"    {x:44s} {n}")

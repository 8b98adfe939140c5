"""Print the host timeline written by `bench.py --trace FILE` (development aid)."""
import json
import sys

ev = json.load(open(sys.argv[1]))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
t00 = min(e[1] for e in ev)
tids = {}
for e in sorted(ev, key=lambda e: e[1]):
    name, t0, t1, tid, extra = e
    k = tids.setdefault(tid, len(tids))
    if (t1 - t0) * 1e3 >= thr or name.startswith("e2e"):
        print("%9.1f %9.1f  T%d %-28s %s" % ((t0 - t00) * 1e3, (t1 - t0) * 1e3, k, name, extra if extra else ""))

import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d["ms_per_step"], d["poisson_ms"], d["value"])

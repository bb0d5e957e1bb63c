import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), d["gpu_launches"], d["counters"])

import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "ms_per_step", round(d["ms_per_step"], 3))

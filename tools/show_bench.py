"""Print the JSON line(s) of a bench.py log compactly: headline keys, then one row per `extra` entry."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path, errors="replace").read().splitlines():
        if line.startswith('{"metric"') or line.startswith('{"impl"'):
            d = json.loads(line)
            x = d.pop("extra", {})
            print(json.dumps(d)[:6000])
            for k, v in x.items():
                for r in v:
                    print("  ", k, r)

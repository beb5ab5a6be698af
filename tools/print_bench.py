"""Print the headline fields of a bench.py JSON line.  usage: python tools/print_bench.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms_per_step %.3f value %.0f e2e %.0f loss %.6f host_enqueue_ms %.2f" % (
    d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss"], d.get("host_enqueue_ms_per_step", -1)))

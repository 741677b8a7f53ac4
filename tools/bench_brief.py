"""Print the headline fields of bench.py JSON lines read from stdin (one per line); extra args are echoed as a label."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(" ".join(sys.argv[1:]), "depth", d.get("run", {}).get("batches_in_flight"), "value", d.get("value"), "e2e", d.get("e2e", {}).get("value"),
          "ms/step", d.get("ms_per_step"), "sched", d.get("schedule"), "pool", d.get("run", {}).get("fano_pool"), "clk", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"), "parity", d.get("parity"), flush=True)

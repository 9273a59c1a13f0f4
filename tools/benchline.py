"""Prints the key figures of a bench.py JSON line read from stdin (last line)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print(tag, "NO JSON LINE")
    sys.exit(1)
d = json.loads(lines[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
c = d.get("clocks") or {}
print(tag, "value=%.0f ms/step=%.4f demod_GBs=%.0f frac=%.3f launch_ms=%.4f share=%.3f sm_mhz=%s e2e=%s launches=%s exact=%s" % (
    d["value"], d["ms_per_step"], r.get("achieved", 0), r.get("frac", 0), r.get("avg_launch_ms", 0),
    r.get("share_of_step", 0), c.get("sm_mhz"), e.get("value"), d.get("gpu_launches"),
    (d.get("config") or {}).get("payloads_exact")))

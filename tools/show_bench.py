"""One-line summary of a bench.py JSON line piped in on stdin (used by the sweep scripts):
    python bench.py ... | python tools/show_bench.py <tag>
For files use tools/bench_summary.py; with a terminal on stdin this refuses instead of waiting forever."""
import json,sys
tag=sys.argv[1]
if sys.stdin.isatty():
    sys.exit("show_bench.py reads a bench.py line from stdin; for files use tools/bench_summary.py")
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
    r=d.get("rollout") or {}
    c=d.get("single_stream_chained") or {}
    e=d.get("e2e") or {}
    print("%s us/step=%.3f frac=%.3f chained_us=%.3f resident_us=%.3f rollout=%.4g rfrac=%.3f e2e=%.3g clk=%s %s" % (tag, d["ms_per_step"]*1e3, d["roofline"]["frac"], c.get("ms_per_step",0)*1e3, d["l2_resident"]["ms_per_step"]*1e3, r.get("value",0), r.get("frac",0), e.get("value",0), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(tag, "FAILED", e)

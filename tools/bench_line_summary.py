"""one-line summary of a bench.py JSON line: python tools/bench_line_summary.py LABEL < line.json   (reads STDIN)"""
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "launch_ms", round(d["roofline"]["launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"]))

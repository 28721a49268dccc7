#!/usr/bin/env python
"""One-line summary of a bench.py JSON line:  python tools/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    print(f, "n_gpus", d.get("n_gpus"), "evals/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4),
          "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d.get("roofline", {}).get("kernel_ms", 0), 4),
          "exchange", d.get("config", {}).get("exchange"), "clocks", d.get("clocks"))

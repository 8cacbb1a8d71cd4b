"""Print the `regimes` entries of bench.py JSON lines: python tools/show_regimes.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(f, "value", round(d.get("value", 0), 1), "ms", round(d.get("ms_per_step", 0), 4))
        for r in d.get("regimes", []):
            ro = r["roofline"]
            ft = r.get("f32_tensor", {})
            print("  B=%4d k=%3d %-6s ms=%.4f qps=%.0f %s frac=%.3f f32rows=%s exec=%s %s reruns=%s/%s | %s" % (
                r["batch"], r["k"], r["regime"], r["ms_per_step"], r["value"], ro["bound"], ro["frac"],
                ("%.3f" % ro["frac_on_fp32_row_bytes"]) if "frac_on_fp32_row_bytes" in ro else "-",
                ("%.3f" % ro["executed_frac"]) if "executed_frac" in ro else "-",
                ft.get("shadow"), ft.get("exact_reruns"), ft.get("queries"), r.get("config", "")[:60]))

"""Render bench.py --kernel-table JSON as a markdown table (profiles/)."""
import json, sys
rows = json.load(open(sys.argv[1]))
peak_tf, peak_gb = 1389.9, 6560.0
tot = sum(r["ms"] for r in rows)
print("| # | kernel | call site | ms | achieved | of measured peak | share |")
print("|---|---|---|---|---|---|---|")
for i, r in enumerate(rows):
    peak = peak_tf if r["unit"] == "flop" else peak_gb
    print("| %d | %s | %s | %.3f | %.1f %s | %.0f %% | %.1f %% |" % (
        i, r["kernel"], r["what"], r["ms"], r["achieved"], r["achieved_unit"], 100 * r["achieved"] / peak, 100 * r["ms"] / tot))
print("\nsum of kernel times %.3f ms per step (CUDA events around each call, steady state, 10 steps)" % tot)

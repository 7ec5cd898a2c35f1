"""Print the per-kernel roofline table of a bench.py JSON line.  python tools/bench_summary.py file.json [n_rows]"""
import json
import sys
line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
d = json.loads(line)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("%s: %.1f %s, %.2f ms/step, e2e %.1f, launches %d, clocks %s" % (d["metric"], d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
for k in ("cpu_baseline", "torch_stock", "ddp_check", "frozen_backbone"):
    if d.get(k):
        print("  %s: %s" % (k, {a: b for a, b in d[k].items() if a not in ("what", "sample")}))
if d.get("infer"):
    print("  infer: %.0f frames/s, one clip %s" % (d["infer"]["value"], d["infer"].get("one_clip_latency")))
r = d["roofline"]
print("  roofline top: %s (%s) frac %.3f share %.3f; eager kernel sum %.2f ms" % (r["kernel"], r["bound"], r["frac"], r["share_of_step"], r["eager_step_ms_sum_of_kernels"]))
print("| share | us/step | calls | bound | frac | kernel family |\n|---|---|---|---|---|---|")
for k in r["kernels"][:n]:
    print("| %4.1f%% | %7.1f | %3.0f | %s | %s | %s |" % (100 * k["share_of_step"], k["us_per_step"], k["calls_per_step"], k["bound"], ("%.3f" % k["frac"]) if k["frac"] is not None else "-", k["kernel"]))

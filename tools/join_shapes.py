"""Join the ncu launch list of tools/profile_step.py with the per-call shape log (gpurun_out/step_shapes.json):
   python tools/join_shapes.py launches.csv step_shapes.json skip_first_n_launches
Prints per distinct (kind, shape): launches, total us, avg us, TFLOP/s (ncu durations)."""
import collections, csv, json, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if l.startswith('"')]))[int(sys.argv[3]):]
shapes = json.load(open(sys.argv[2]))
tc_conv = [float(r["Metric Value"]) / 1e3 for r in rows if "conv_tc_kernel" in r["Kernel Name"]]
tc_wg = [float(r["Metric Value"]) / 1e3 for r in rows if "wgrad_tc_kernel" in r["Kernel Name"]]
sc = [s for s in shapes if s["engine"] == "tcgen05" and s["kind"] != "wgrad"]
sw = [s for s in shapes if s["engine"] == "tcgen05" and s["kind"] == "wgrad"]
assert len(sc) <= len(tc_conv) and len(sw) <= len(tc_wg), (len(sc), len(tc_conv), len(sw), len(tc_wg))
tc_conv, tc_wg = tc_conv[-len(sc):], tc_wg[-len(sw):]   # the logged step is the last one
agg = collections.OrderedDict()
for s, t in list(zip(sc, tc_conv)) + list(zip(sw, tc_wg)):
    key = (s["kind"],) + tuple(s["shape"])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += s["flops"]
tot = sum(a[1] for a in agg.values())
print(f"tcgen05 launches {len(sc) + len(sw)}  total {tot/1e3:.3f} ms")
print("kind  (B, phases, t_src, t_dst, c_src, c_dst, k, dil, stride, groups)            n   total_us  avg_us  TFLOP/s")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:5s} {str(k[1:]):62s} {a[0]:3d} {a[1]:9.1f} {a[1]/a[0]:7.1f} {a[2]/a[1]/1e6:8.1f}")

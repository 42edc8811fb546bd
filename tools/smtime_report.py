"""SM-time accounting of tools/timeline.py output: for every kernel of one step, CTAs x duration / 148 (a kernel of 20 CTAs
that runs 10 us holds 20 SMs for 10 us, not the chip).   python tools/smtime_report.py gpurun_out/timeline.json"""
import json, collections, re, sys
rows = json.load(open(sys.argv[1]))
step = rows[len(rows) // 2:]
def short(nm):
    m = re.search(r'(\w+)(<|\()', nm.replace('void ', '').replace('stg::(anonymous namespace)::', '').replace('stg::', ''))
    return (m.group(1) if m else nm)[:28]
t0 = step[0]['ts']; end = max(r['ts'] + r['dur'] for r in step) - t0
sm = collections.Counter(); busy = collections.Counter(); cnt = collections.Counter(); hist = collections.Counter()
for r in step:
    g = r.get('grid') or [1, 1, 1]
    ctas = g[0] * g[1] * g[2]
    n = short(r['name'])
    sm[n] += min(ctas, 148) * r['dur'] / 148.0; busy[n] += r['dur']; cnt[n] += 1
    if n in ('conv_tc_kernel', 'wgrad_tc_kernel'):
        hist[(n, '<=37' if ctas <= 37 else '<=74' if ctas <= 74 else '<=111' if ctas <= 111 else '<148' if ctas < 148 else '148')] += r['dur']
print(f"span {end:.0f} us; sum of SM-time {sum(sm.values()):.0f} us-chip ({sum(sm.values()) / end:.2f} of the chip held)")
print("| kernel | launches | busy us | SM-time us-chip |\n|---|---:|---:|---:|")
for k, v in sm.most_common(16): print(f"| {k} | {cnt[k]} | {busy[k]:.0f} | {v:.0f} |")
print("busy us of the tcgen05 kernels by CTA count:", {f"{k[0][:5]} {k[1]}": round(v) for k, v in sorted(hist.items())})

"""Summarise tools/timeline.py output: per-kernel busy time, time by concurrency level, solo time by kernel and window.
   python tools/timeline_report.py gpurun_out/timeline.json [window_us]"""
import json, collections, re, sys
rows = json.load(open(sys.argv[1])); win_us = int(sys.argv[2]) if len(sys.argv) > 2 else 250
step = rows[len(rows) // 2:]
t0 = step[0]['ts']
def short(nm):
    m = re.search(r'(\w+)(<|\()', nm.replace('void ', '').replace('stg::(anonymous namespace)::', '').replace('stg::', ''))
    return (m.group(1) if m else nm)[:28]
for r in step: r['ts'] -= t0; r['n'] = short(r['name'])
end = max(r['ts'] + r['dur'] for r in step)
pts = []
for i, r in enumerate(step):
    pts.append((r['ts'], 1, i)); pts.append((r['ts'] + r['dur'], -1, i))
pts.sort(key=lambda x: (x[0], x[1]))
active = set(); last = 0; hist = collections.Counter(); solo = collections.Counter(); segs = []
for t, d, i in pts:
    dt = t - last
    hist[min(len(active), 6)] += dt
    if len(active) == 1 and dt > 0:
        j = next(iter(active)); solo[step[j]['n']] += dt; segs.append((last, dt, step[j]['n']))
    last = t
    if d == 1: active.add(i)
    else: active.discard(i)
print(f"events {len(step)}  span {end:.0f} us")
print("time by concurrency level (us):", {k: round(v) for k, v in sorted(hist.items())})
bt = collections.Counter(); cnt = collections.Counter()
for r in step: bt[r['n']] += r['dur']; cnt[r['n']] += 1
print(f"total kernel time {sum(bt.values()):.0f} us")
print("| kernel | launches | busy us | solo us |\n|---|---:|---:|---:|")
for k, v in bt.most_common(24): print(f"| {k} | {cnt[k]} | {v:.0f} | {solo.get(k, 0):.0f} |")
win = collections.defaultdict(collections.Counter)
for s, dt, nm in segs: win[int(s // win_us)][nm] += dt
print("\nsolo time by window:")
for w in sorted(win): print(f"{w * win_us:6d} us  " + ", ".join(f"{k} {v:.0f}" for k, v in win[w].most_common(4)))

"""Timeline of one traced solve (ILQG_TRACE=<file>): per pass, when each launch site finished on its stream.
usage: python tools/trace_view.py trace.txt [first_pass] [passes]"""
import sys
NAMES = {0: "pass begins", 1: "K_lq", 2: "K_bwd", 13: "first-window rollout", 14: "first-window merit", 15: "first-window decide",
         23: "tier rollout", 24: "tier merit", 25: "tier decide", 6: "K_lq (queue list)", 3: "prologue rollout", 4: "prologue merit"}
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 3
count = int(sys.argv[3]) if len(sys.argv) > 3 else 2
starts = [i for i, (tag, _) in enumerate(rows) if tag == 0]
for p in range(first, min(first + count, len(starts))):
    lo, hi = starts[p], starts[p + 1] if p + 1 < len(starts) else len(rows)
    t0 = rows[lo][1]
    print(f"--- pass {p} (us after its first stamp; duration {((rows[hi][1] if hi < len(rows) else rows[-1][1]) - t0) / 1e3:.1f} us)")
    last = {0: t0, 1: None}
    for tag, t in rows[lo:hi]:
        side = tag % 10 if False else tag % 2 if False else (tag % 10)
        site, stream = tag // 10, tag % 10
        prev = last[stream] if last[stream] is not None else t
        print(f"  {'side' if stream else 'main'}  {(t - t0) / 1e3:9.1f}  (+{(t - prev) / 1e3:7.1f})  {NAMES.get(site, site)}")
        last[stream] = t

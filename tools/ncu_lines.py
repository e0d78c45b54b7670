"""Per-source-line aggregation of an `ncu --page source --print-source cuda,sass --csv` export: instructions executed and
warp-stall samples per CUDA source line (python tools/ncu_lines.py <csv> [top N] [substring of the kernel name])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else None
hdr, cur_file, agg, use = None, None, {}, True
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        use = only is None or only in r[1]
        continue
    if not use:
        continue
    if r[0] == 'Line No':
        hdr = r
        ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= ii or r[2] != '-':
        continue
    try:
        ins, sm = float(r[ii]), float(r[si])
    except ValueError:
        continue
    a = agg.setdefault((cur_file, int(r[0]), r[1]), [0, 0])
    a[0] += ins
    a[1] += sm
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print('total warp instructions %.4g, samples %d' % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0] - kv[1][1] * ti / ts)[:top]:
    print('%-16s %5d inst %5.1f%% samp %5.1f%%  %s' % (k[0][:16], k[1], 100 * v[0] / ti, 100 * v[1] / ts, k[2][:105]))

"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/r02_launches.csv -> profiles/r02_launches.md)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0].replace('void ', '')
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1; a[1] += float(r[vi].replace(',', ''))
total = sum(a[1] for a in agg.values())
print('# Launch list of one timed step (bench.py --steps 1 under `ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none`)\n')
print('One step = one whole calibration of 512 samples (16 batches x 32 images, both phases).  ncu serialises launches and runs them cold, so only the')
print('SHARES are comparable with the live CUDA-event timing in the bench line (`roofline.ms_per_launch` x 16 / `ms_per_step`).\n')
print('| kernel | launches | total us | share | grid | block |\n|---|---|---|---|---|---|')
for k, a in agg.items():
    print(f'| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / total:.1f} % | {a[2]} | {a[3]} |')
print(f'| **all** | {sum(a[0] for a in agg.values())} | {total / 1e3:.1f} | | | |')

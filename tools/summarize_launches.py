"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel name launches, summed device time and DRAM traffic of the LAST `--last` launches (= one step of
`bench.py --steps 1 --warmup 1 --no-graph`; the launches before it are the autotune pass and the warm-up step).

    python tools/summarize_launches.py gpurun_out/launches.csv --last 811 --json profiles/r01_traffic_pix2pix.json
"""
import argparse
import csv
import json
import re
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('catb::', '')
    return re.sub(r'void ', '', name).strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--last', type=int, required=True)
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    per_id = OrderedDict()
    for r in rd:
        if 'Metric Name' not in r or not r.get('ID', '').isdigit():
            continue
        e = per_id.setdefault(int(r['ID']), {'name': short(r['Kernel Name'])})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m.startswith('gpu__time_duration'):
            e['us'] = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
        elif m.startswith('dram__bytes'):
            e[m.split('.')[0]] = v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
    launches = list(per_id.values())[-a.last:]
    agg = OrderedDict()
    for e in launches:
        g = agg.setdefault(e['name'], {'launches': 0, 'us': 0.0, 'dram_read': 0.0, 'dram_write': 0.0})
        g['launches'] += 1
        g['us'] += e.get('us', 0.0)
        g['dram_read'] += e.get('dram__bytes_read', 0.0)
        g['dram_write'] += e.get('dram__bytes_write', 0.0)
    total = sum(g['us'] for g in agg.values())
    print(f'{len(launches)} launches, {total / 1e3:.2f} ms summed (cold cache, serialised: compare shares, not absolutes)')
    print(f'{"kernel":44s} {"launches":>8s} {"time us":>10s} {"share":>7s} {"DRAM MB":>10s} {"GB/s":>8s}')
    for name, g in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        mb = (g['dram_read'] + g['dram_write']) / 1e6
        print(f'{name[:44]:44s} {g["launches"]:8d} {g["us"]:10.0f} {100 * g["us"] / total:6.1f}% {mb:10.1f} {mb / max(g["us"], 1e-9) * 1e3:8.0f}')
    if a.json:
        out = {'launches': len(launches), 'summed_ms': total / 1e3,
               'kernels': {n: dict(g, share=g['us'] / total, traffic_bytes_per_launch=(g['dram_read'] + g['dram_write']) / g['launches'])
                           for n, g in agg.items()}}
        with open(a.json, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()

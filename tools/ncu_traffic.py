#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --page raw --csv` dump of the three k_front launches at the full bench
size: dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by frame size.
    ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv <source note>"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn, rd, wr = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}
for r in rows[2:]:
    m = re.search(r'k_front(?:_pair)?<\(?(?:int\))?(\d+)', r[kn])
    if m:
        out[m.group(1)] = float(r[rd].replace(',', '')) * scale[units[rd]] + float(r[wr].replace(',', '')) * scale[units[wr]]
print(json.dumps(out, indent=1))

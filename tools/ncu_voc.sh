#!/bin/bash
# ncu full capture of the late (wide-time, few-channel) vocoder conv launches at bench size; only small CSVs are kept
set -u
mkdir -p gpurun_out
O=gpurun_out
R=/tmp/r1_voc_late_full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:voc_conv_mma_kernel -s 83 -c 8 -f -o $R \
  python tools/profile_frame.py --frames 256 --vocoder > $O/r1_voc_late_full.log 2>&1
ncu -i $R.ncu-rep --page raw --csv > $O/r1_voc_late_full.raw.csv 2>/dev/null
ncu -i $R.ncu-rep --page source --csv --launch-skip 7 --launch-count 1 > /tmp/src.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/src.csv')))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
keep = ['Address', 'Source', '# Samples', 'Instructions Executed'] + [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
with open('gpurun_out/r1_voc_late_k7_source.csv', 'w', newline='') as f:
    w = csv.writer(f); w.writerow(keep)
    for r in rows[2:]:
        if len(r) == len(hdr): w.writerow([r[ix[k]] for k in keep])
PY
ls -la $O | tail -4

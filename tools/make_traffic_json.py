"""profiles/fdmt_traffic.json from an ncu dram__bytes capture of one bfFdmtExecute
(tools/gpu_round.sh).  bench.py reports it as roofline.traffic."""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, im, iv, iid = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
per = {}
for r in rows[1:]:
    per.setdefault((int(r[iid]), r[ik][:60]), {})[r[im]] = float(r[iv].replace(',', ''))
kern = [dict(kernel=k[1], dram_read=v.get('dram__bytes_read.sum'), dram_write=v.get('dram__bytes_write.sum'),
             duration_ns=v.get('gpu__time_duration.sum')) for k, v in sorted(per.items())]
total = sum((k['dram_read'] or 0) + (k['dram_write'] or 0) for k in kern)
json.dump(dict(source=src.split('/')[-1], what='dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE '
               'bfFdmtExecute on the bench workload (ncu, not a timing)', dram_bytes_per_call=total, kernels=kern),
          open(dst, 'w'), indent=1)
print(total)

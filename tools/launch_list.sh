#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of the library's own kernels in a short bench run:
#   tools/launch_list.sh OUT.csv [bench args...]     (prints a per-kernel summary of the LAST step)
out=$1; shift
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:^(k[0-9a]|ka_|philox)' -c 80 --csv --log-file "$out" \
    python bench.py --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 0 "$@" > gpurun_out/launch_list_bench.log 2>&1
python - "$out" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rows[1:]:
    agg.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0][:60]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
items = sorted(agg.items())
names = [k[1] for k, _ in items]
# the step that holds the longest launch of the first kernel (the bench steps; the e2e leg's calls are smaller)
first = names[0]
starts = [i for i, n in enumerate(names) if n == first] + [len(items)]
best = max(range(len(starts) - 1), key=lambda j: sum(v["gpu__time_duration.sum"] for _, v in items[starts[j]:starts[j + 1]]))
print(f"{'kernel':62s} {'us':>10s} {'dram MB':>10s}")
for (i, n), v in items[starts[best]:starts[best + 1]]:
    print(f"{n:62s} {v['gpu__time_duration.sum'] / 1e3:10.1f} {(v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / 1e6:10.1f}")
PY

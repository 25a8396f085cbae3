#!/usr/bin/env python
"""profiles/k2_pipes.json from one `ncu --set full` capture of k2_dirichlet_kernel and the JSON line bench.py
printed in that same (profiled) run - bench.py copies these figures into `roofline.k2`.

    python tools/k2_pipes.py gpurun_out/r2_k2.ncu-rep gpurun_out/ncu_k2.log > profiles/k2_pipes.json

instr_per_draw = thread instructions executed by the captured launch / gamma draws of that launch
(images per launch x pairs per image x T x C_out, all taken from the bench line)."""
import csv
import io
import json
import subprocess
import sys

rep, log = sys.argv[1], sys.argv[2]
line = None
for ln in open(log):
    if ln.startswith("{") and '"metric"' in ln:
        line = json.loads(ln)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
r = rows[2]
g = lambda k: float(r[ix[k]].replace(",", ""))
cfg = line["config"]
draws = cfg["batch_per_gpu"] * line["roofline"]["k2"]["draws_per_image"]
tinst = g("smsp__inst_executed.sum") * g("smsp__thread_inst_executed_per_inst_executed.ratio")
print(json.dumps(dict(
    instr_per_draw=tinst / draws,
    issue_active=g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
    pipe_xu=g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active") / 100,
    pipe_alu=g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100,
    pipe_fma=g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100,
    pipe_lsu=g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active") / 100,
    warps_active=g("sm__warps_active.avg.pct_of_peak_sustained_active") / 100,
    registers=int(g("launch__registers_per_thread")),
    ncu_source=f"{rep.split('/')[-1]}: k2_dirichlet_kernel<false>, {cfg['batch_per_gpu']} images per launch, "
               f"{g('gpu__time_duration.sum'):.3f} ms under ncu"), indent=1))

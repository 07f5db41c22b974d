"""Runs ON THE GPU BOX: one `ncu --set full` capture of every kernel of ONE training step of the bench workload (after warm-up), reduced to
a per-launch CSV (duration, DRAM bytes, L2->SM bytes, tensor-pipe %, DRAM %, registers, shared memory, grid) that fits the gpurun copy-back
limit.  The .ncu-rep itself (hundreds of MB) stays on the box.

    python scripts/ncu_step_summary.py [tf32|3xtf32] [out.csv] [light]
`light`: collect only the metrics of the summary (a handful of replay passes per kernel instead of the ~40 of `--set full`; the 3xTF32 step, 36 ms of kernels
over multi-GB working sets, does not finish a full-set capture within minutes).
The step runs eagerly (plan cache off) inside a cudaProfilerStart / Stop range after three warm-up steps."""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "ncu_step_%s.csv" % mode)
rep = "/tmp/ncu_step_%s" % mode
env = dict(os.environ, AGX_PLAN_CACHE="0")
light = len(sys.argv) > 3 and sys.argv[3] == "light"
METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,"
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,"
           "lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active")
cmd = ["ncu"] + (["--metrics", METRICS] if light else ["--set", "full"]) + ["--clock-control", "none", "--profile-from-start", "off", "-f", "-o", rep,
       sys.executable, os.path.join(ROOT, "scripts", "one_step.py"), mode, "profile"]
subprocess.run(cmd, check=True, env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_size"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow(["id"] + ["%s [%s]" % (hdr[i], units[i]) if units[i] else hdr[i] for i in idx])
    for k, r in enumerate(rows[2:]):
        vals = [r[i] for i in idx]
        vals[0] = vals[0].replace("void ", "").split("(")[0][:90]
        wr.writerow([k] + vals)
print("wrote", out, len(rows) - 2, "launches")

cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3_t_all.log
tail -3 gpurun_out/r3_t_all.log
timeout 420 python bench.py --steps 10 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r3_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["modes"], d["micro"].get("gemm_8192_3xtf32"), d["parity"]["3xtf32"], d["clocks"])
P

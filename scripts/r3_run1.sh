set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or matmul or conv" 2>&1 | tail -15 > gpurun_out/r3_t1.log
tail -5 gpurun_out/r3_t1.log
for lib in base new; do
  if [ $lib = base ]; then export AGB200_LIB=$GRAFT_REPO_ROOT/rust-autograd_b200/lib/libagb200_base.so; else unset AGB200_LIB; fi
  echo "== $lib" >> gpurun_out/r3_gemm3x.txt; timeout 120 python scripts/bench_gemm3x.py >> gpurun_out/r3_gemm3x.txt 2>&1
  echo "== $lib" >> gpurun_out/r3_conv3x.txt; timeout 120 python scripts/bench_conv3x.py 0 >> gpurun_out/r3_conv3x.txt 2>&1
done
cat gpurun_out/r3_gemm3x.txt gpurun_out/r3_conv3x.txt

#!/bin/bash
# round-2 GPU session 2: squaring block size x instruction cache (4 builds), ncu --set full of the three dominant
# kernels on the best build, bench at N = 10^6 with the online verifier
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
df -h /dev/shm > gpurun_out/s2_shm.txt
for bs in 96 48 32 16; do
  VMX_LIBRARY_PATH=$PWD/build/variants/libvmx_bs$bs.so timeout 300 python tools/prof_driver.py 100000 3072 quick > gpurun_out/s2_prof_bs$bs.log 2>&1
done
grep -H "expMulExp\|exp_scalar (3071\|exp_var (613\|exp_fixed" gpurun_out/s2_prof_bs*.log
best=$(python - <<'PY'
import re,glob
best=None
for f in sorted(glob.glob("gpurun_out/s2_prof_bs*.log")):
    m=re.search(r"expMulExp.*?([0-9.]+) ms", open(f).read())
    if m and (best is None or float(m.group(1))<best[0]): best=(float(m.group(1)), re.search(r"bs(\d+)", f).group(1))
print(best[1] if best else "16")
PY
)
echo "best variant: bs$best" | tee gpurun_out/s2_best.txt
export VMX_LIBRARY_PATH=$PWD/build/variants/libvmx_bs$best.so
timeout 300 python tools/prof_driver.py 1000000 3072 > gpurun_out/s2_prof_1m.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu --no-other --phases > gpurun_out/s2_bench_1m.log 2> gpurun_out/s2_bench_1m.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu --no-other --offline-verify --e2e-steps 3 > gpurun_out/s2_bench_1m_offline.log 2> gpurun_out/s2_bench_1m_offline.err; echo "bench offline rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_exp_fixed|k_exp_var2' -s 2 -c 2 -o gpurun_out/s2_prof_fixed_var2 python tools/ncu_driver_all.py 100000 > gpurun_out/s2_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_seg_prod' -s 0 -c 2 -o gpurun_out/s2_prof_segprod python tools/ncu_driver_all.py 100000 > gpurun_out/s2_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep

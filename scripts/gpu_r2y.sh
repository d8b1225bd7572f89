mkdir -p gpurun_out
timeout 600 ncu --target-processes all --set full --clock-control none --import-source on -k regex:k_bam_inflate -s 1 -c 1 -f -o gpurun_out/r02_inflate python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --file-reads 12000000 > gpurun_out/ncu_inflate.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_inflate.log | cut -c1-300

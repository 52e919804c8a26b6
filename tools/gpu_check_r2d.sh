# third gpurun call of the round: the dual-sim controller GPU tests again (sanity bound fixed, contact queries added) and one
# `ncu --set full` capture of rg_step_kernel for the final build (same command as the launch list of the first call)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rearrange_arm.py -m gpu -q -p no:cacheprovider > gpurun_out/r2d_gputests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2d_gputests.log
tail -15 gpurun_out/r2d_gputests.log
RG_CPU_BASELINE_SECONDS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:rg_step_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2c python bench.py --steps 2 --warmup 3 > gpurun_out/r2d_ncu.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/

# second gpurun call of the round: the GPU tests added after the first one (dual-sim controller, heterogeneous scenes) and the
# bench line of the dual-simulation rearrange loop
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_rearrange_arm.py tests/test_rearrange_scene.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c_gputests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2c_gputests.log
tail -25 gpurun_out/r2c_gputests.log
timeout 300 python bench.py --config rearrange_blocks_tcp --steps 40 --warmup 10 > gpurun_out/r2c_bench_rearrange_blocks_tcp.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench_rearrange_blocks_tcp.json; tail -5 gpurun_out/r2c_bench.err

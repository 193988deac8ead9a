mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_loader_gpu.py tests/test_real_data.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/h2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h2_tests.log
timeout 200 python profiles/micro/fuel_property.py --sets cn ysi --epochs 400 > gpurun_out/fuel.json 2> gpurun_out/fuel.err; rc=$?; echo "fuel rc=$rc" >> gpurun_out/fuel.err
if [ $rc -ne 0 ]; then CUDA_LAUNCH_BLOCKING=1 timeout 100 python profiles/micro/fuel_property.py --sets cn --epochs 4 > gpurun_out/fuel_blk.json 2> gpurun_out/fuel_blk.err; echo "blk rc=$?" >> gpurun_out/fuel_blk.err; fi
tail -3 gpurun_out/h2_tests.log; tail -2 gpurun_out/fuel.err

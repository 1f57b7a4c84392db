python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1.log 2>&1
python tools/frame_timeline.py c4_cloud_100k 240 256 > gpurun_out/r2_timeline_flush.log 2>&1
python tools/frame_timeline.py c4_cloud_100k 240 0 > gpurun_out/r2_timeline_noflush.log 2>&1
SWEEP_WARM=150 python tools/perf_sweep.py > gpurun_out/r2_sweep_t2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitizer_subject.py > gpurun_out/r2_san_memcheck_1gpu.log 2>&1
timeout 420 compute-sanitizer --tool racecheck python tools/sanitizer_subject.py > gpurun_out/r2_san_racecheck_1gpu.log 2>&1
timeout 240 compute-sanitizer --tool synccheck python tools/sanitizer_subject.py > gpurun_out/r2_san_synccheck_1gpu.log 2>&1
tail -3 gpurun_out/r2_pytest_1.log gpurun_out/r2_sweep_t2.log gpurun_out/r2_san_*.log

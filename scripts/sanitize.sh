#!/bin/bash
# compute-sanitizer over the GPU parity tests (run on the GPU box: gpurun -- bash scripts/sanitize.sh).
# memcheck on the whole -m gpu suite of the optimizer / ESDF / multi paths; racecheck on a reduced selection (it is ~50x slower).
set -u
mkdir -p gpurun_out
export ALORE_SANITIZE_SMALL=1
compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_r02.log \
  python -m pytest tests/test_optimizer_gpu.py tests/test_esdf_gpu.py tests/test_golden.py -x -q -m gpu \
  -k "not division and not 4096 and not 8192 and not 2048 and not long_trajectories" > gpurun_out/memcheck_r02.pytest.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck_r02.pytest.txt
compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/racecheck_r02.log \
  python -m pytest tests/test_optimizer_gpu.py -x -q -m gpu -k "cost_and_gradient_batch or small_piece_counts or penalty_batch or final_collision or config1" \
  > gpurun_out/racecheck_r02.pytest.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/racecheck_r02.pytest.txt
ALORE_OPT_WAVE=1 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_wave_r02.log \
  python -m pytest tests/test_optimizer_gpu.py -x -q -m gpu -k "config1 or batch_of_legs or collision_replans" > gpurun_out/memcheck_wave_r02.pytest.txt 2>&1
echo "memcheck(wave) rc=$?" >> gpurun_out/memcheck_wave_r02.pytest.txt
tail -n 3 gpurun_out/memcheck_r02.pytest.txt gpurun_out/racecheck_r02.pytest.txt gpurun_out/memcheck_wave_r02.pytest.txt
grep -h "ERROR SUMMARY" gpurun_out/memcheck_r02.log gpurun_out/racecheck_r02.log gpurun_out/memcheck_wave_r02.log
# ESDF far-field path (K2 masks/flags, conditional K1b, K2e band envelope, warp-per-cell quirk column)
compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_esdf_r02.log \
  python -m pytest tests/test_esdf_gpu.py tests/test_alt_paths_gpu.py -x -q -m gpu -k "(small_maps or far_field or sub_window or named_maps) and not 4096 and not 8192" \
  > gpurun_out/memcheck_esdf_r02.pytest.txt 2>&1
echo "memcheck(esdf) rc=$?" >> gpurun_out/memcheck_esdf_r02.pytest.txt
compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/racecheck_esdf_r02.log \
  python -m pytest tests/test_alt_paths_gpu.py -x -q -m gpu -k "far_field and (knobs0 or knobs2)" \
  > gpurun_out/racecheck_esdf_r02.pytest.txt 2>&1
echo "racecheck(esdf) rc=$?" >> gpurun_out/racecheck_esdf_r02.pytest.txt
tail -n 3 gpurun_out/memcheck_esdf_r02.pytest.txt gpurun_out/racecheck_esdf_r02.pytest.txt
grep -h "ERROR SUMMARY" gpurun_out/memcheck_esdf_r02.log gpurun_out/racecheck_esdf_r02.log

set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py tests/test_gpu_tsdf.py tests/test_gpu_mesh.py -q -p no:cacheprovider -x -k "edge_sizes or binned_sort or bit_exact or mesh_matches or tma_staging or optional_input" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py tests/test_gpu_mesh.py -q -p no:cacheprovider -x -k "edge_sizes or mesh_matches or tma_staging" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_racecheck.log

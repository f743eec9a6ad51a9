# full GPU verification: tests, smoke, bench (both arms). usage: gpu_verify.sh <tag>
TAG=${1:-dev}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_gputest.log
python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/${TAG}_gputest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_gputest.log

timeout 600 python -m pytest tests/test_peer_exchange.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -25
SUFFIX=_peer timeout 300 bash tools/gpu_scale.sh 2 --e2e-rounds 5
SUFFIX=_c2_peer timeout 300 bash tools/gpu_scale.sh 2 --workload c2_bunny_chess --e2e-rounds 0
SUFFIX=_c2_nccl timeout 300 bash tools/gpu_scale.sh 2 --workload c2_bunny_chess --e2e-rounds 0 --collective native

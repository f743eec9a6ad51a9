N=${1:-8}
SUFFIX=_c2_peer2 timeout 300 bash tools/gpu_scale.sh $N --workload c2_bunny_chess --e2e-rounds 0
SUFFIX=_c1_peer2 timeout 300 bash tools/gpu_scale.sh $N --workload c1_spheres_bezier --e2e-rounds 0
SUFFIX=_peer2 timeout 600 bash tools/gpu_scale.sh $N --e2e-rounds 0
SUFFIX=_nccl2 timeout 600 bash tools/gpu_scale.sh $N --e2e-rounds 0 --collective native
SUFFIX=_strong_peer2 timeout 300 bash tools/gpu_scale.sh $N --photons $((16777216 / N)) --e2e-rounds 0

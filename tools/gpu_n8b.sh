N=${1:-8}
SUFFIX=_peer timeout 600 bash tools/gpu_scale.sh $N
SUFFIX=_c1_peer timeout 300 bash tools/gpu_scale.sh $N --workload c1_spheres_bezier --e2e-rounds 0
SUFFIX=_c2_peer timeout 300 bash tools/gpu_scale.sh $N --workload c2_bunny_chess --e2e-rounds 0
SUFFIX=_c2_nccl timeout 300 bash tools/gpu_scale.sh $N --workload c2_bunny_chess --e2e-rounds 0 --collective native
SUFFIX=_strong_peer timeout 300 bash tools/gpu_scale.sh $N --photons $((16777216 / N)) --e2e-rounds 0
SUFFIX=_c5_peer timeout 300 bash tools/gpu_scale.sh $N --workload c5_dragon_4096 --e2e-rounds 0 --steps 2 --warmup 1
cd cgraytracing_b200/host && for how in peer nccl; do timeout 120 ./example_main bunny 320 240 400000 4 /tmp/o_$how.ppm ../assets $N $how | tail -1; done; timeout 120 ./example_main bunny 320 240 400000 4 /tmp/o_1.ppm ../assets 1 | tail -1

# variant benches, then the GPU tests with one variant as the library. usage: VARIANTS="A B" TESTLIB=B bash tools/gpu_tv2.sh
bash tools/gpu_variants.sh
if [ -n "$TESTLIB" ]; then CGRT_LIB=$PWD/gpurun_variants_$TESTLIB.so python -m pytest tests -m gpu -x -q 2>&1 | tail -6; fi

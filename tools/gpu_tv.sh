# tests (fail-fast) then variant benches. usage: VARIANTS="A B" bash tools/gpu_tv.sh [pytest -k expr]
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -12
bash tools/gpu_variants.sh

timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tier_splits" 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -30
bash tools/run_gpu_variants.sh "nobar_c1" "nobar_c2" "nobar_c3" "nobar_c4" "nobar_c3 -- --config c3 --steps 5" "nobar -- --config c3 --steps 5"

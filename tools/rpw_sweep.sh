for r in 2 4 8 16 32; do echo "KC_RPW=$r"; KC_RPW=$r python tools/firstbench.py 2>&1 | grep -E "float32 B=4096|float64 B=4096"; done

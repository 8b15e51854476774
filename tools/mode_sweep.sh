for m in narrow wide; do echo "MODE=$m"; KC_ROLLOUT_MODE=$m python tools/firstbench.py 2>&1 | grep -E "B=4096|B=16384"; done

#!/bin/bash
# development run for the tensor-core march kernels: each stage under its own timeout so a hang costs little
mkdir -p gpurun_out
python -c "import __graft_entry__" 2>&1 | tail -1
echo "== selftest" ; timeout -s KILL 240 python -m pytest tests/test_gpu_knode_tc.py -q -k "selftest" -p no:cacheprovider 2>&1 | tail -15
echo "== fwd golden" ; timeout -s KILL 240 python -m pytest tests/test_gpu_knode_tc.py -q -k "vs_reference or ragged or matches_simt" -p no:cacheprovider 2>&1 | tail -40
echo "== bwd" ; timeout -s KILL 300 python -m pytest tests/test_gpu_knode_tc.py -q -k "bptt or trains" -p no:cacheprovider 2>&1 | tail -40

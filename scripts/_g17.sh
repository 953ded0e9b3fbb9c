set -u
export TFQB_PEER_TIMEOUT_S=5
TFQB_JIT_MIN_AMPS=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sharded_state_peer_memory_exchange" 2>&1 | grep -v "^$" | tail -25
echo "---- old geometry"
TFQB_JIT_FWD_SEQ=1 TFQB_JIT_MIN_AMPS=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sharded_state_peer_memory_exchange" 2>&1 | tail -4

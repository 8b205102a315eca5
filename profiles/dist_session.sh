#!/bin/bash
# One multi-GPU box session: the multi-GPU parity tests, then bench.py at N ranks with the NVSwitch multicast
# exchange (default) and with unicast pushes (A/B).
#   gpurun --gpus N --timeout 900 -- 'bash profiles/dist_session.sh N r2'
set -u
N=${1:-2}
TAG=${2:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${TAG}_n$N.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py -x -q -m gpu ) > $OUT/pytest_dist_${TAG}_n$N.log 2>&1
tail -4 $OUT/pytest_dist_${TAG}_n$N.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 \
      > $OUT/bench_${TAG}_n${N}_$name.json 2> $OUT/bench_${TAG}_n${N}_$name.err
  echo "$name rc=$?"; tail -c 600 $OUT/bench_${TAG}_n${N}_$name.json
}
run mc KGAT_PEER_MULTICAST=1
run unicast KGAT_PEER_MULTICAST=0
ls -la $OUT | tail -12

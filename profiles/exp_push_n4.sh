# push bandwidth of one rank, 4 ranks pushing at once: multicast vs unicast (profiles/push_bw.py)
set -u
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
KGAT_PEER_MULTICAST=1 timeout 200 $TR profiles/push_bw.py > $OUT/push_bw_n4_mc.json 2> $OUT/push_bw_n4_mc.err
KGAT_PEER_MULTICAST=0 timeout 200 $TR profiles/push_bw.py > $OUT/push_bw_n4_uc.json 2> $OUT/push_bw_n4_uc.err
cat $OUT/push_bw_n4_mc.json $OUT/push_bw_n4_uc.json; tail -3 $OUT/push_bw_n4_mc.err

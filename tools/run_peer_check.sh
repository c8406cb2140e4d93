# 2-GPU A/B of the panel broadcast paths of DistChol (tools/dist_check.py): NCCL vs fused TRSM -> peer stores
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29512"
N=${N:-40960}; T=${T:-1024}
for m in on off; do
echo "== n=$N peer $m"; timeout 400 $TR tools/dist_check.py --size $N --tile $T --reps 3 --peer $m 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-1200
done
echo "== n=$N p2p"; LGP_DIST_MULTIMEM=0 timeout 400 $TR tools/dist_check.py --size $N --tile $T --reps 3 --peer on 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-1200

# 8-GPU checks of the fused panel path: oracle parity at n=3000 (2x2 on 4 ranks, 2x4 on 8), then config C5 (n=150000) A/B
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
echo "== 4 ranks 2x2 peer on"; timeout 200 $TR4 tools/dist_check.py --size 3000 --tile 256 --grid 2x2 --oracle --peer on 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-700
echo "== 8 ranks 2x4 peer on"; timeout 200 $TR8 tools/dist_check.py --size 3000 --tile 256 --grid 2x4 --oracle --peer on 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-700
echo "== 8 ranks 2x4 p2p"; LGP_DIST_MULTIMEM=0 timeout 200 $TR8 tools/dist_check.py --size 3000 --tile 256 --grid 2x4 --oracle --peer on 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-700
echo "== C5 n=150000 peer on"; timeout 400 $TR8 tools/dist_check.py --size 150000 --tile 1024 --reps 2 --peer on --out gpurun_out/dist8_n150k_T1024_peer.json 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-1200
echo "== C5 n=150000 peer off"; timeout 400 $TR8 tools/dist_check.py --size 150000 --tile 1024 --reps 2 --peer off --out gpurun_out/dist8_n150k_T1024_nccl.json 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-1200

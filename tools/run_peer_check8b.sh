TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
echo "== 8 ranks 2x4 peer on (oracle, n=3000)"; timeout 200 $TR8 tools/dist_check.py --size 3000 --tile 256 --grid 2x4 --oracle --peer on 2>&1 | grep -E "DIST_CHECK|Error|Traceback" | cut -c1-300
echo "== C5 n=150000 peer on"; timeout 400 $TR8 tools/dist_check.py --size 150000 --tile 1024 --reps 2 --peer on --out gpurun_out/dist8_n150k_T1024_peer2.json 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | cut -c1-1200

TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
echo "== multimem 2x1"; timeout 200 $TR tools/dist_check.py --size 3000 --tile 256 --grid 2x1 --oracle --peer on 2>&1 | grep -E "DIST_CHECK|Error|error|Traceback" | cut -c1-300
echo "== p2p 2x1"; LGP_DIST_MULTIMEM=0 timeout 200 $TR tools/dist_check.py --size 3000 --tile 256 --grid 2x1 --oracle --peer on 2>&1 | grep -E "DIST_CHECK|Error|error|Traceback" | cut -c1-300
echo "== n=40960 multimem"; timeout 300 $TR tools/dist_check.py --size 40960 --tile 1024 --reps 3 --peer on 2>&1 | grep -E "^\{|DIST_CHECK|Error|Traceback" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['peer_mode'], d['factor_tflops'], d['resid'])
    else: print(l.strip())"

for d in 0 16 23; do echo "DBG=$d"; LDCONV_DBG=$d python benchmarks/ldconv_layers.py --iters 3 2>/dev/null | grep offset_conv_fwd | grep tcgen05 | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    if r['layer'] in (15,10,1): print('  L',r['layer'],r['us'])
"; done

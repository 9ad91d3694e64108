for r in 0 8 12 16 20 24 32 48; do echo "rows=$r"; LDC_MARCH_ROWS=$r python tools/kernel_sweep.py --kernels residual --variants 2 --sizes 256,1024,2048,4096 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print('   n=%d %.5f ms %.1f%%' % (d['n'], d['ms'], 100*d['frac']))
    except Exception: pass"; done

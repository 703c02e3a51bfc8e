for v in "$@"; do for fm in 5 3; do
echo -n "$v fused=$fm: "; SPDY_LIB=$PWD/gpurun_in/lib_$v.so SPDY_FUSED=$fm python tools/time_classes.py 512 2>&1 | tail -1 | python -c "
import sys,ast
l=sys.stdin.read().strip()
try:
    d=ast.literal_eval(l); print('fft_inv',d['fft_inv'],'total',round(sum(d.values()),3))
except Exception as e: print(l[-300:])
"
done; done

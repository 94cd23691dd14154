#!/bin/bash
for f in tools/tmp/variants/lib_w4_b1.so tools/tmp/variants/lib_w4_b6.so tools/tmp/variants/lib_w4_b8.so tools/tmp/variants/lib_w8_b1.so tools/tmp/variants/lib_w8_b3.so tools/tmp/variants/lib_w8_b4.so tools/tmp/variants/lib_w2_b10.so; do
DYNENV_B200_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'env-steps/s', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],2))"
done

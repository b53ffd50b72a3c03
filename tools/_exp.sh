set -e
timeout 200 python tools/conv_tune.py 6 8192 400
TZ_LIB=$PWD/takzero_b200/build/variants/nob.so timeout 200 python tools/conv_tune.py 6 8192 400
timeout 200 python tools/conv_tune.py 6 8192 400
